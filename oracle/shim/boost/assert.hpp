// empty Boost stub (TEST INFRASTRUCTURE): included by KalmanFilter.h:6-17, nothing from it is used
#pragma once
