#include "../pcl/shim_pcl.h"
