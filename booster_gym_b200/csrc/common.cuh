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
