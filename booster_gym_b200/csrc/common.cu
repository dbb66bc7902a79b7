// common.cu - last-error string, version and ABI self-check of libb200t1.so
#include <string.h>

#include "common.cuh"

namespace b200 {
static thread_local char g_err[512] = "";
int set_error(int code, const char* msg) {
    snprintf(g_err, sizeof g_err, "%s", msg);
    return code;
}
int set_cuda_error(cudaError_t e, const char* where) {
    snprintf(g_err, sizeof g_err, "CUDA error %d (%s) at %s", (int)e, cudaGetErrorString(e), where);
    return B200_ERR_CUDA;
}
}  // namespace b200

// FP32 FMA peak of this device, measured: 8 independent FMA chains per thread (latency 4 x 2 issue slots covered with 8 warps per
// scheduler), 148 x 8 CTAs of 256 threads.  The denominator of k_physics' roofline (SURVEY 8d asks for a measured figure).
__global__ void __launch_bounds__(256) k_fma_peak(float* out, int iters, float a, float b) {
    float x0 = threadIdx.x, x1 = x0 + 1.f, x2 = x0 + 2.f, x3 = x0 + 3.f, x4 = x0 + 4.f, x5 = x0 + 5.f, x6 = x0 + 6.f, x7 = x0 + 7.f;
#pragma unroll 1
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int k = 0; k < 16; ++k) {
            x0 = fmaf(x0, a, b); x1 = fmaf(x1, a, b); x2 = fmaf(x2, a, b); x3 = fmaf(x3, a, b);
            x4 = fmaf(x4, a, b); x5 = fmaf(x5, a, b); x6 = fmaf(x6, a, b); x7 = fmaf(x7, a, b);
        }
    }
    const float s = ((x0 + x1) + (x2 + x3)) + ((x4 + x5) + (x6 + x7));
    if (s == 123.456f) out[0] = s;   // never true: keeps the chains alive
}

extern "C" {
int b200_fma_peak(double* tflops, void* stream) {
    if (!tflops) return b200::set_error(B200_ERR_ARG, "b200_fma_peak: null pointer");
    cudaStream_t st = (cudaStream_t)stream;
    int sms = 148;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    float* out = nullptr;
    CUDA_TRY(cudaMalloc(&out, sizeof(float)));
    cudaEvent_t e0, e1;
    CUDA_TRY(cudaEventCreate(&e0));
    CUDA_TRY(cudaEventCreate(&e1));
    const int iters = 4096, grid = sms * 8;
    k_fma_peak<<<grid, 256, 0, st>>>(out, 64, 0.999f, 0.001f);   // warm-up
    double best = 0.0;
    for (int rep = 0; rep < 5; ++rep) {
        CUDA_TRY(cudaEventRecord(e0, st));
        k_fma_peak<<<grid, 256, 0, st>>>(out, iters, 0.999f, 0.001f);
        CUDA_TRY(cudaEventRecord(e1, st));
        CUDA_TRY(cudaEventSynchronize(e1));
        float ms = 0.f;
        CUDA_TRY(cudaEventElapsedTime(&ms, e0, e1));
        const double fl = 2.0 * 8 * 16 * (double)iters * 256.0 * grid;
        const double tf = fl / (ms * 1e-3) / 1e12;
        best = tf > best ? tf : best;
    }
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    cudaFree(out);
    *tflops = best;
    return B200_OK;
}
const char* b200_last_error(void) { return b200::g_err; }
int b200_version(void) { return 100; }
int b200_sizeof(int what) {
    switch (what) {
        case 0: return (int)sizeof(B200T1ModelF);
        case 1: return (int)sizeof(B200T1Config);
        case 2: return (int)sizeof(B200PpoConfig);
        case 3: return (int)sizeof(B200T1ModelD);
    }
    return -1;
}
}
