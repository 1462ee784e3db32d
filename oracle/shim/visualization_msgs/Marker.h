#include "../ros/ros.h"
