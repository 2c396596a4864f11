// stub (TEST INFRASTRUCTURE): see kfshim_msgs.h
#pragma once
#include "kfshim_msgs.h"
