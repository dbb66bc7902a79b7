// common.cuh - error plumbing shared by the translation units of libb200t1.so
#pragma once
#include <cuda_runtime.h>
#include <stdio.h>

#include "../../include/b200_t1.h"

#ifndef PHYS_BLOCK
#define PHYS_BLOCK 32
#endif

namespace b200 {
int set_error(int code, const char* msg);
int set_cuda_error(cudaError_t e, const char* where);
// cudaFuncAttributeMaxDynamicSharedMemorySize is a per-DEVICE attribute: remember which devices a kernel has been configured on
// (a process may create learners / envs on several devices)
template <typename K> inline cudaError_t ensure_dynamic_smem(K kernel, int bytes, unsigned long long& device_mask) {
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return e;
    if (dev < 64 && ((device_mask >> dev) & 1ull)) return cudaSuccess;
    e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
    if (e == cudaSuccess && dev < 64) device_mask |= 1ull << dev;
    return e;
}
inline int launch_status(const char* what) {
    cudaError_t e = cudaPeekAtLastError();
    if (e != cudaSuccess) { cudaGetLastError(); return set_cuda_error(e, what); }
    return B200_OK;
}
}  // namespace b200

#define CUDA_TRY(expr)                                                        \
    do {                                                                      \
        cudaError_t _e = (expr);                                              \
        if (_e != cudaSuccess) return b200::set_cuda_error(_e, #expr);        \
    } while (0)
