#include "shim_pcl.h"
