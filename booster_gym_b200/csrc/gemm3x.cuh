// gemm3x.cuh - fp32-accurate tensor-core GEMM for the actor/critic MLP layers (utils/model.py:9-26).
//
// The reference runs these contractions as cuBLAS SGEMM in fp32 and the parity bar is 1e-5 relative, which a single
// TF32/BF16 pass (~1e-3) cannot meet.  Every product is therefore evaluated as an error-compensated split
// a = a_hi + a_lo (both TF32):  a*b ~= a_lo*b_hi + a_hi*b_lo + a_hi*b_hi, accumulated in fp32 ("3xTF32"), on
// mma.sync.m16n8k8.tf32 tiles.  One kernel template covers the three contractions of a layer:
//   forward   Y[M,N]  = act(X[M,K] W[N,K]^T + b)            A row-major,   B "col-major" (both reduction-contiguous)
//   dgrad     dX[M,K] = (dY[M,N] W[N,K]) * ELU'(H[M,K])      A row-major,   B row-major
//   wgrad     dW[N,K] += dY[M,N]^T X[M,K]  (split over M)    A col-major,   B row-major, atomic accumulate
// CTA tile 128 x 128 x 32, 8 warps (2 x 4), warp tile 64 x 32, register-prefetch double buffering through shared memory.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace b200 {

enum GemmEpilogue { EPI_BIAS_ELU = 0, EPI_BIAS = 1, EPI_ELU_GRAD = 2, EPI_ATOMIC = 3 };

struct GemmArgs {
    const float* A;
    const float* B;
    float* C;
    const float* bias;  // [Cn] or null            (EPI_BIAS*)
    const float* aux;   // post-activation H [I, ldaux] (EPI_ELU_GRAD)
    int I, Cn, R;       // output rows, output cols, reduction length
    int lda, ldb, ldc, ldaux;
    int cn_store;       // columns actually stored (<= Cn), e.g. 47 of a 48-wide padded tile
    int r_chunk;        // reduction rows per CTA along blockIdx.z (split-R); 0 = whole R
    int r_valid_b;      // reduction indices >= this read B as 0 (K = 47/61 of a padded X); <=0 = R
    float* C_hi;        // optional (EPI_BIAS_ELU): tf32 hi / lo halves of the stored value, same layout as C - the operand
    float* C_lo;        //                          format of the tcgen05 kernels (gemm_tc.cuh)
};

#ifndef G_ACC_SPLIT
#define G_ACC_SPLIT 1
#endif
#define G_BM 128
#define G_BN 128
#define G_BK 32
#define G_THREADS 256
#define G_LDS_R (G_BK + 4)   // [i][r] layout stride  (36: fragment reads hit 32 distinct banks)
#define G_LDS_K (G_BM + 8)   // [r][i] layout stride  (136)
#define G_TILE_FLOATS (G_BM * G_LDS_R)  // 4608 >= 32 * 136 = 4352
#define G_SMEM_BYTES (4 * G_TILE_FLOATS * (int)sizeof(float))

// round-to-nearest (ties away) to tf32 on the bit pattern: 2 ALU ops; cvt.rna.tf32.f32 compiles to ~6 on sm_100 (finite inputs only)
__device__ __forceinline__ uint32_t f2tf32(float x) { return (__float_as_uint(x) + 0x1000u) & 0xFFFFE000u; }
__device__ __forceinline__ void split_tf32(float x, uint32_t& hi, uint32_t& lo) {
    hi = f2tf32(x);
    lo = f2tf32(x - __uint_as_float(hi));
}
__device__ __forceinline__ void mma_tf32(float* d, const uint32_t* a, const uint32_t* b) {
    asm volatile(
        "mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
        : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
        : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
}

// 4 consecutive floats starting at p (element index `first` of a run with `avail` valid elements left); zero fill
__device__ __forceinline__ float4 load4_guard(const float* p, int avail, bool vec_ok) {
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (avail >= 4 && vec_ok) {
        v = *reinterpret_cast<const float4*>(p);
    } else {
        if (avail > 0) v.x = p[0];
        if (avail > 1) v.y = p[1];
        if (avail > 2) v.z = p[2];
        if (avail > 3) v.w = p[3];
    }
    return v;
}

template <bool A_T, bool B_T, int EPI>
__global__ void __launch_bounds__(G_THREADS) k_gemm3x(const GemmArgs g) {
    extern __shared__ __align__(16) float smem[];
    float* sA[2] = {smem, smem + G_TILE_FLOATS};
    float* sB[2] = {smem + 2 * G_TILE_FLOATS, smem + 3 * G_TILE_FLOATS};
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int gq = lane >> 2, tq = lane & 3;
    const int warp_m = warp >> 2, warp_n = warp & 3;  // 2 x 4 warps, warp tile 64 x 32
    const int i0 = blockIdx.x * G_BM, c0 = blockIdx.y * G_BN;
    int r_begin = 0, r_end = g.R;
    if (g.r_chunk > 0) {
        r_begin = blockIdx.z * g.r_chunk;
        r_end = min(g.R, r_begin + g.r_chunk);
        if (r_begin >= r_end) return;
    }
    const int r_valid_b = (g.r_valid_b > 0) ? g.r_valid_b : g.R;
    const bool vecA = ((g.lda & 3) == 0) && ((reinterpret_cast<uintptr_t>(g.A) & 15) == 0);
    const bool vecB = ((g.ldb & 3) == 0) && ((reinterpret_cast<uintptr_t>(g.B) & 15) == 0);
    const int nk = (r_end - r_begin + G_BK - 1) / G_BK;

    float4 ra[4], rb[4];
    auto load_tiles = [&](int kt) {
        const int rk = r_begin + kt * G_BK;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int idx = tid + G_THREADS * j;
            if (!A_T) {  // A[i][r], float4 along r
                const int row = idx >> 3, r4 = (idx & 7) * 4;
                const int i = i0 + row, r = rk + r4;
                ra[j] = (i < g.I) ? load4_guard(g.A + (size_t)i * g.lda + r, r_end - r, vecA) : make_float4(0, 0, 0, 0);
            } else {     // A[r][i], float4 along i
                const int rr = idx >> 5, i4 = (idx & 31) * 4;
                const int r = rk + rr, i = i0 + i4;
                ra[j] = (r < r_end) ? load4_guard(g.A + (size_t)r * g.lda + i, g.I - i, vecA) : make_float4(0, 0, 0, 0);
            }
            if (!B_T) {  // B stored [c][r] (W[N,K]), float4 along r
                const int col = idx >> 3, r4 = (idx & 7) * 4;
                const int c = c0 + col, r = rk + r4;
                rb[j] = (c < g.Cn) ? load4_guard(g.B + (size_t)c * g.ldb + r, min(r_end, r_valid_b) - r, vecB) : make_float4(0, 0, 0, 0);
            } else {     // B stored [r][c], float4 along c
                const int rr = idx >> 5, c4 = (idx & 31) * 4;
                const int r = rk + rr, c = c0 + c4;
                rb[j] = (r < r_end && r < r_valid_b) ? load4_guard(g.B + (size_t)r * g.ldb + c, g.Cn - c, vecB) : make_float4(0, 0, 0, 0);
            }
        }
    };
    auto store_tiles = [&](int buf) {
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int idx = tid + G_THREADS * j;
            if (!A_T) *reinterpret_cast<float4*>(sA[buf] + (idx >> 3) * G_LDS_R + (idx & 7) * 4) = ra[j];
            else *reinterpret_cast<float4*>(sA[buf] + (idx >> 5) * G_LDS_K + (idx & 31) * 4) = ra[j];
            if (!B_T) *reinterpret_cast<float4*>(sB[buf] + (idx >> 3) * G_LDS_R + (idx & 7) * 4) = rb[j];
            else *reinterpret_cast<float4*>(sB[buf] + (idx >> 5) * G_LDS_K + (idx & 31) * 4) = rb[j];
        }
    };

    float acc[4][4][4];
#pragma unroll
    for (int mi = 0; mi < 4; ++mi)
#pragma unroll
        for (int ni = 0; ni < 4; ++ni)
#pragma unroll
            for (int q = 0; q < 4; ++q) acc[mi][ni][q] = 0.f;

    load_tiles(0);
    store_tiles(0);
    __syncthreads();
    for (int kt = 0; kt < nk; ++kt) {
        const int buf = kt & 1;
        if (kt + 1 < nk) load_tiles(kt + 1);
        const float* As = sA[buf];
        const float* Bs = sB[buf];
#pragma unroll
        for (int kb = 0; kb < G_BK; kb += 8) {
            uint32_t ah[4][4], al[4][4], bh[4][2], bl[4][2];
#pragma unroll
            for (int mi = 0; mi < 4; ++mi) {
                const int rbase = warp_m * 64 + mi * 16;
                float x0, x1, x2, x3;
                if (!A_T) {
                    x0 = As[(rbase + gq) * G_LDS_R + kb + tq];
                    x1 = As[(rbase + gq + 8) * G_LDS_R + kb + tq];
                    x2 = As[(rbase + gq) * G_LDS_R + kb + tq + 4];
                    x3 = As[(rbase + gq + 8) * G_LDS_R + kb + tq + 4];
                } else {
                    x0 = As[(kb + tq) * G_LDS_K + rbase + gq];
                    x1 = As[(kb + tq) * G_LDS_K + rbase + gq + 8];
                    x2 = As[(kb + tq + 4) * G_LDS_K + rbase + gq];
                    x3 = As[(kb + tq + 4) * G_LDS_K + rbase + gq + 8];
                }
                split_tf32(x0, ah[mi][0], al[mi][0]);
                split_tf32(x1, ah[mi][1], al[mi][1]);
                split_tf32(x2, ah[mi][2], al[mi][2]);
                split_tf32(x3, ah[mi][3], al[mi][3]);
            }
#pragma unroll
            for (int ni = 0; ni < 4; ++ni) {
                const int cbase = warp_n * 32 + ni * 8;
                float y0, y1;
                if (!B_T) {
                    y0 = Bs[(cbase + gq) * G_LDS_R + kb + tq];
                    y1 = Bs[(cbase + gq) * G_LDS_R + kb + tq + 4];
                } else {
                    y0 = Bs[(kb + tq) * G_LDS_K + cbase + gq];
                    y1 = Bs[(kb + tq + 4) * G_LDS_K + cbase + gq];
                }
                split_tf32(y0, bh[ni][0], bl[ni][0]);
                split_tf32(y1, bh[ni][1], bl[ni][1]);
            }
#pragma unroll
            for (int mi = 0; mi < 4; ++mi)
#pragma unroll
                for (int ni = 0; ni < 4; ++ni) {
#if G_ACC_SPLIT
                    // the tensor core adds into its fp32 accumulator with truncation; keeping each k-step's partial sum
                    // in a fresh accumulator and adding it with a round-to-nearest FADD removes the bias that otherwise
                    // grows linearly with the reduction length (measured: 8x the fp32 SGEMM error at K = 256)
                    float t[4] = {0.f, 0.f, 0.f, 0.f};
                    mma_tf32(t, al[mi], bh[ni]);
                    mma_tf32(t, ah[mi], bl[ni]);
                    mma_tf32(t, ah[mi], bh[ni]);
#pragma unroll
                    for (int q = 0; q < 4; ++q) acc[mi][ni][q] += t[q];
#else
                    mma_tf32(acc[mi][ni], al[mi], bh[ni]);
                    mma_tf32(acc[mi][ni], ah[mi], bl[ni]);
                    mma_tf32(acc[mi][ni], ah[mi], bh[ni]);
#endif
                }
        }
        if (kt + 1 < nk) store_tiles(buf ^ 1);
        __syncthreads();
    }

    // ---- epilogue -------------------------------------------------------------------------------------------------
#pragma unroll
    for (int mi = 0; mi < 4; ++mi) {
#pragma unroll
        for (int ni = 0; ni < 4; ++ni) {
            const int c = c0 + warp_n * 32 + ni * 8 + 2 * tq;
#pragma unroll
            for (int half = 0; half < 2; ++half) {
                const int i = i0 + warp_m * 64 + mi * 16 + gq + 8 * half;
                if (i >= g.I) continue;
                float v0 = acc[mi][ni][2 * half], v1 = acc[mi][ni][2 * half + 1];
                if (EPI == EPI_ATOMIC) {
                    if (c < g.cn_store) atomicAdd(g.C + (size_t)i * g.ldc + c, v0);
                    if (c + 1 < g.cn_store) atomicAdd(g.C + (size_t)i * g.ldc + c + 1, v1);
                } else {
                    if (EPI == EPI_BIAS_ELU || EPI == EPI_BIAS) {
                        if (g.bias) {
                            if (c < g.Cn) v0 += g.bias[c];
                            if (c + 1 < g.Cn) v1 += g.bias[c + 1];
                        }
                        if (EPI == EPI_BIAS_ELU) {
                            v0 = (v0 > 0.f) ? v0 : expm1f(v0);
                            v1 = (v1 > 0.f) ? v1 : expm1f(v1);
                        }
                    } else if (EPI == EPI_ELU_GRAD) {
                        if (c < g.cn_store) { const float h = g.aux[(size_t)i * g.ldaux + c]; v0 *= (h > 0.f) ? 1.f : (h + 1.f); }
                        if (c + 1 < g.cn_store) { const float h = g.aux[(size_t)i * g.ldaux + c + 1]; v1 *= (h > 0.f) ? 1.f : (h + 1.f); }
                    }
                    if (EPI == EPI_BIAS_ELU && g.C_hi) {
                        const float h0 = __uint_as_float(f2tf32(v0)), h1 = __uint_as_float(f2tf32(v1));
                        if (c < g.cn_store) { g.C_hi[(size_t)i * g.ldc + c] = h0; g.C_lo[(size_t)i * g.ldc + c] = __uint_as_float(f2tf32(v0 - h0)); }
                        if (c + 1 < g.cn_store) { g.C_hi[(size_t)i * g.ldc + c + 1] = h1; g.C_lo[(size_t)i * g.ldc + c + 1] = __uint_as_float(f2tf32(v1 - h1)); }
                    }
                    if (c + 1 < g.cn_store && ((g.ldc & 1) == 0)) {
                        *reinterpret_cast<float2*>(g.C + (size_t)i * g.ldc + c) = make_float2(v0, v1);
                    } else {
                        if (c < g.cn_store) g.C[(size_t)i * g.ldc + c] = v0;
                        if (c + 1 < g.cn_store) g.C[(size_t)i * g.ldc + c + 1] = v1;
                    }
                }
            }
        }
    }
}

template <bool A_T, bool B_T, int EPI>
inline cudaError_t launch_gemm3x(const GemmArgs& g, int split, cudaStream_t st) {
    static bool configured = false;
    if (!configured) {
        cudaError_t e = cudaFuncSetAttribute(k_gemm3x<A_T, B_T, EPI>, cudaFuncAttributeMaxDynamicSharedMemorySize, G_SMEM_BYTES);
        if (e != cudaSuccess) return e;
        configured = true;
    }
    dim3 grid((g.I + G_BM - 1) / G_BM, (g.Cn + G_BN - 1) / G_BN, split > 0 ? split : 1);
    k_gemm3x<A_T, B_T, EPI><<<grid, G_THREADS, G_SMEM_BYTES, st>>>(g);
    return cudaPeekAtLastError();
}

}  // namespace b200
