// TEST INFRASTRUCTURE: read_xml lives in the ptree stub
#pragma once
#include "ptree.hpp"
