// Force-included (-include) in front of the reference translation units built into
// oracle/_ref.  The reference's CMake passes the OpenMP thread count as the string
// macro MP_PROC_NUM="<cores-4>" (R/CMakeLists.txt:26-31); here the macro expands to a
// call so that one build can be timed at several thread counts.
#pragma once
#ifdef __cplusplus
extern "C" const char *ref_mp_proc_num(void);
#endif
