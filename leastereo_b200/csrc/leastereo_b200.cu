// leastereo_b200 - CUDA translation unit for the CUDA-core kernels and the common C-ABI plumbing.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 (see __graft_entry__.build()).
#include "lea_common.h"
#include <cstdarg>
#include <cstdio>

static thread_local char g_lea_err[512] = "";
void lea_set_error(const char* fmt, ...) {
    va_list ap; va_start(ap, fmt); vsnprintf(g_lea_err, sizeof(g_lea_err), fmt, ap); va_end(ap);
}
extern "C" const char* lea_last_error(void) { return g_lea_err; }
extern "C" int lea_abi_version(void) { return LEA_ABI_VERSION; }
extern "C" int lea_is_device_build(void) { return 1; }

static int lea_post_launch(const char* what) {
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) { lea_set_error("%s: launch failed: %s", what, cudaGetErrorString(e)); return 2; }
    return 0;
}
#define LEA_POST_LAUNCH() lea_post_launch(__func__)

#include "lea_simt_kernels.cuh"
#include "lea_train_kernels.cuh"
#include "lea_io_kernels.cuh"
#include "lea_api_simt.inl"
#include "lea_api_train.inl"
#include "lea_api_io.inl"
