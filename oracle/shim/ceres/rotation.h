// ceres is included by R/include/desc/STDesc.h but nothing on the path uses it
