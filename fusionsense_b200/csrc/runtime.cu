// runtime.cu — library-wide bookkeeping and the small stream-ordered upload used by captured steps.
#include "common.cuh"

#include <string.h>

unsigned long long g_fsb_launches = 0;

// kernels launched by this library since load (monotone counter)
FSB_API uint64_t fsb_launch_count(void) { return __atomic_load_n(&g_fsb_launches, __ATOMIC_RELAXED); }

FSB_API int fsb_abi_version(void) { return 3; }

namespace {

constexpr int UPLOAD_WORDS = 64;  // 256 bytes of kernel arguments

struct UploadArgs {
    uint32_t w[UPLOAD_WORDS];
};

__global__ void upload_small_kernel(UploadArgs a, int n_words, uint32_t* __restrict__ dst) {
    const int i = threadIdx.x;
    if (i < n_words) dst[i] = a.w[i];
}

}  // namespace

FSB_API int fsb_upload_small_max(void) { return UPLOAD_WORDS * 4; }

// The payload rides in the launch arguments (copied by the driver when the launch is enqueued), so the host
// buffer is free again on return and the write lands in stream order: the per-replay inputs of a CUDA-graph
// captured step (camera index, Adam step sizes) need neither pinned staging rings nor a synchronisation.
FSB_API int fsb_upload_small(void* dst_dev, const void* src_host, int bytes, void* stream) {
    if (!dst_dev || !src_host || bytes <= 0 || bytes > UPLOAD_WORDS * 4 || (bytes & 3) ||
        ((uintptr_t)dst_dev & 3))
        return FSB_E_ARG;
    UploadArgs a;
    memset(&a, 0, sizeof(a));
    memcpy(a.w, src_host, (size_t)bytes);
    upload_small_kernel<<<1, UPLOAD_WORDS, 0, (cudaStream_t)stream>>>(a, bytes / 4, (uint32_t*)dst_dev);
    FSB_LAUNCH_CHECK();
    return 0;
}
