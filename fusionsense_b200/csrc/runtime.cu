// runtime.cu — library-wide bookkeeping (no kernels).
#include "common.cuh"

unsigned long long g_fsb_launches = 0;

// kernels launched by this library since load (monotone counter)
FSB_API uint64_t fsb_launch_count(void) { return __atomic_load_n(&g_fsb_launches, __ATOMIC_RELAXED); }

FSB_API int fsb_abi_version(void) { return 1; }
