// tc_proto.cu - standalone prototype of the tcgen05 3xTF32 GEMM (K-major x K-major): C[M,N] = A[M,K] B[N,K]^T
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o tc_proto tc_proto.cu -lcuda
#include <cuda.h>
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <vector>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(1); } } while (0)

static constexpr int BM = 128, BK = 32;

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* b, uint32_t cnt) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(b)), "r"(cnt)); }
__device__ __forceinline__ void mbar_expect_tx(uint64_t* b, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(b)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* b, uint32_t parity) {
    uint32_t done;
    const uint32_t a = smem_u32(b);
    do {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(done) : "r"(a), "r"(parity) : "memory");
    } while (!done);
}
__device__ __forceinline__ void tma_load_2d(const CUtensorMap* map, uint64_t* bar, void* dst, int c0, int c1) {
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
                 ::"r"(smem_u32(dst)), "l"((uint64_t)map), "r"(smem_u32(bar)), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ uint64_t make_desc_k(uint32_t saddr) {  // K-major, SWIZZLE_128B, 8-row groups 1024 B apart
    uint64_t d = 0;
    d |= (uint64_t)((saddr >> 4) & 0x3FFF);
    d |= (uint64_t)1 << 16;             // leading byte offset (unused for swizzled K-major), 16 B
    d |= (uint64_t)(1024 >> 4) << 32;   // stride byte offset
    d |= (uint64_t)1 << 46;             // version = 1 (Blackwell)
    d |= (uint64_t)2 << 61;             // SWIZZLE_128B
    return d;
}
__device__ __forceinline__ void umma_tf32(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t acc) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(tmem_d), "l"(da), "l"(db), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}

template <int BN, int STAGES>
__global__ void __launch_bounds__(192, 1) k_tc(const __grid_constant__ CUtensorMap mAh, const __grid_constant__ CUtensorMap mAl,
                                                const __grid_constant__ CUtensorMap mBh, const __grid_constant__ CUtensorMap mBl,
                                                float* __restrict__ C, int M, int N, int K, int do_store = 1) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    constexpr int A_BYTES = BM * BK * 4, B_BYTES = BN * BK * 4, STAGE_BYTES = 2 * A_BYTES + 2 * B_BYTES;
    uint64_t* full = (uint64_t*)(smem + STAGES * STAGE_BYTES);
    uint64_t* empty = full + STAGES;
    uint64_t* tfull = empty + STAGES;
    uint32_t* tmem_slot = (uint32_t*)(tfull + 1);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int m0 = blockIdx.x * BM, n0 = blockIdx.y * BN;
    const int nk = (K + BK - 1) / BK;

    if (warp == 0 && lane == 0) {
        for (int s = 0; s < STAGES; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
        mbar_init(tfull, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"((uint32_t)BN));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    asm volatile("tcgen05.fence::before_thread_sync;");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;");
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        if (lane == 0) {
            for (int kb = 0; kb < nk; ++kb) {
                const int s = kb % STAGES, ph = (kb / STAGES) & 1;
                mbar_wait(&empty[s], ph ^ 1);
                uint8_t* st = smem + s * STAGE_BYTES;
                mbar_expect_tx(&full[s], STAGE_BYTES);
                tma_load_2d(&mAh, &full[s], st, kb * BK, m0);
                tma_load_2d(&mAl, &full[s], st + A_BYTES, kb * BK, m0);
                tma_load_2d(&mBh, &full[s], st + 2 * A_BYTES, kb * BK, n0);
                tma_load_2d(&mBl, &full[s], st + 2 * A_BYTES + B_BYTES, kb * BK, n0);
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            // instruction descriptor: c=f32 (1<<4), a=tf32 (2<<7), b=tf32 (2<<10), K-major both, n>>3 at 17, m>>4 at 24
            const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);
            for (int kb = 0; kb < nk; ++kb) {
                const int s = kb % STAGES, ph = (kb / STAGES) & 1;
                mbar_wait(&full[s], ph);
                asm volatile("tcgen05.fence::after_thread_sync;");
                const uint32_t a_hi = smem_u32(smem + s * STAGE_BYTES), a_lo = a_hi + A_BYTES, b_hi = a_hi + 2 * A_BYTES, b_lo = b_hi + B_BYTES;
#pragma unroll
                for (int k = 0; k < BK / 8; ++k) {
                    const uint32_t off = k * 32;  // 8 tf32 = 32 bytes inside the 128-byte swizzle row
                    umma_tf32(tmem_base, make_desc_k(a_lo + off), make_desc_k(b_hi + off), idesc, (kb | k) ? 1u : 0u);
                    umma_tf32(tmem_base, make_desc_k(a_hi + off), make_desc_k(b_lo + off), idesc, 1u);
                    umma_tf32(tmem_base, make_desc_k(a_hi + off), make_desc_k(b_hi + off), idesc, 1u);
                }
                umma_commit(&empty[s]);
            }
            umma_commit(tfull);
        }
    } else {
        const int q = warp & 3;  // TMEM lane quarter this warp may access
        mbar_wait(tfull, 0);
        asm volatile("tcgen05.fence::after_thread_sync;");
        const int row = m0 + q * 32 + lane;
#pragma unroll 1
        for (int c = 0; c < BN / 32; ++c) {
            uint32_t r[32];
            const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(c * 32);
            asm volatile(
                "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
                : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
                  "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]),
                  "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]),
                  "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
                : "r"(taddr));
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
            if (row < M && (do_store || r[5] == 0x12345u)) {
                for (int j = 0; j < 32; ++j) {
                    const int col = n0 + c * 32 + j;
                    if (col < N) C[(size_t)row * N + col] = __uint_as_float(r[j]);
                }
            }
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;");
    __syncthreads();
    if (warp == 1) {
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)BN));
    }
}


// ---- wgrad prototype: D[NO, KI] = sum_m dY[m, NO]^T X[m, KI]  (both operands MN-major: the reduction index m is the row index)
__device__ __forceinline__ uint64_t make_desc_mn(uint32_t saddr) {  // MN-major, SWIZZLE_128B: 32-float MN chunks 4096 B apart, 8-row k groups 1024 B apart
    uint64_t d = 0;
    d |= (uint64_t)((saddr >> 4) & 0x3FFF);
    d |= (uint64_t)(4096 >> 4) << 16;   // leading byte offset: stride between 32-float MN chunks (one TMA box)
    d |= (uint64_t)(512 >> 4) << 32;    // stride byte offset: stride between 4-row k groups
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)1 << 61;             // SWIZZLE_128B_BASE32B: the only MN-major layout for 32-bit operands
    return d;
}
template <int BN, int STAGES>
__global__ void __launch_bounds__(192, 1) k_tc_wgrad(const __grid_constant__ CUtensorMap mYh, const __grid_constant__ CUtensorMap mYl,
                                                      const __grid_constant__ CUtensorMap mXh, const __grid_constant__ CUtensorMap mXl,
                                                      float* __restrict__ D, int M, int NO, int KI, int chunk) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    constexpr int A_BYTES = BM * BK * 4, B_BYTES = BN * BK * 4, STAGE_BYTES = 2 * A_BYTES + 2 * B_BYTES;
    uint64_t* full = (uint64_t*)(smem + STAGES * STAGE_BYTES);
    uint64_t* empty = full + STAGES;
    uint64_t* tfull = empty + STAGES;
    uint32_t* tmem_slot = (uint32_t*)(tfull + 1);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int r_begin = blockIdx.x * chunk, r_end = min(M, r_begin + chunk);
    const int n0 = blockIdx.y * BM, k0 = blockIdx.z * BN;
    const int nk = (r_end - r_begin + BK - 1) / BK;
    if (warp == 0 && lane == 0) {
        for (int s = 0; s < STAGES; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
        mbar_init(tfull, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"((uint32_t)BN));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    asm volatile("tcgen05.fence::before_thread_sync;");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;");
    const uint32_t tmem_base = *tmem_slot;
    if (warp == 0) {
        if (lane == 0) {
            for (int kb = 0; kb < nk; ++kb) {
                const int s = kb % STAGES, ph = (kb / STAGES) & 1;
                mbar_wait(&empty[s], ph ^ 1);
                uint8_t* st = smem + s * STAGE_BYTES;
                mbar_expect_tx(&full[s], STAGE_BYTES);
                const int r = r_begin + kb * BK;
                for (int c = 0; c < BM / 32; ++c) {
                    tma_load_2d(&mYh, &full[s], st + c * 4096, n0 + c * 32, r);
                    tma_load_2d(&mYl, &full[s], st + A_BYTES + c * 4096, n0 + c * 32, r);
                }
                for (int c = 0; c < BN / 32; ++c) {
                    tma_load_2d(&mXh, &full[s], st + 2 * A_BYTES + c * 4096, k0 + c * 32, r);
                    tma_load_2d(&mXl, &full[s], st + 2 * A_BYTES + B_BYTES + c * 4096, k0 + c * 32, r);
                }
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | (1u << 15) | (1u << 16) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);
            for (int kb = 0; kb < nk; ++kb) {
                const int s = kb % STAGES, ph = (kb / STAGES) & 1;
                mbar_wait(&full[s], ph);
                asm volatile("tcgen05.fence::after_thread_sync;");
                const uint32_t a_hi = smem_u32(smem + s * STAGE_BYTES), a_lo = a_hi + A_BYTES, b_hi = a_hi + 2 * A_BYTES, b_lo = b_hi + B_BYTES;
#pragma unroll
                for (int k = 0; k < BK / 8; ++k) {
                    const uint32_t off = k * 1024;  // 8 samples = one 8-row swizzle group
                    umma_tf32(tmem_base, make_desc_mn(a_lo + off), make_desc_mn(b_hi + off), idesc, (kb | k) ? 1u : 0u);
                    umma_tf32(tmem_base, make_desc_mn(a_hi + off), make_desc_mn(b_lo + off), idesc, 1u);
                    umma_tf32(tmem_base, make_desc_mn(a_hi + off), make_desc_mn(b_hi + off), idesc, 1u);
                }
                umma_commit(&empty[s]);
            }
            umma_commit(tfull);
        }
    } else {
        const int q = warp & 3;
        mbar_wait(tfull, 0);
        asm volatile("tcgen05.fence::after_thread_sync;");
        const int row = n0 + q * 32 + lane;
#pragma unroll 1
        for (int c = 0; c < BN / 32; ++c) {
            uint32_t r[32];
            const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(c * 32);
            asm volatile(
                "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
                : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
                  "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]),
                  "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]),
                  "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
                : "r"(taddr));
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
            if (row < NO) {
                for (int j = 0; j < 32; ++j) {
                    const int col = k0 + c * 32 + j;
                    if (col < KI) atomicAdd(&D[(size_t)row * KI + col], __uint_as_float(r[j]));
                }
            }
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;");
    __syncthreads();
    if (warp == 1) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)BN));
}

typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                             const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeFn get_encode() {
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult q;
    CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q));
    return (EncodeFn)fn;
}
static CUtensorMap make_map(EncodeFn enc, float* base, int rows, int cols, int box_rows, CUtensorMapSwizzle sw = CU_TENSOR_MAP_SWIZZLE_128B) {
    CUtensorMap m;
    cuuint64_t gdim[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
    cuuint64_t gstr[1] = {(cuuint64_t)cols * 4};
    cuuint32_t box[2] = {32, (cuuint32_t)box_rows};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = enc(&m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, base, gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     sw, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { printf("cuTensorMapEncodeTiled failed %d\n", (int)r); exit(1); }
    return m;
}
static float tf32_rna(float x) {
    uint32_t u; memcpy(&u, &x, 4);
    u = (u + 0x1000u) & 0xFFFFE000u;
    float y; memcpy(&y, &u, 4);
    return y;
}

template <int BN, int STAGES> static void run(int M, int N, int K) {
    EncodeFn enc = get_encode();
    std::vector<float> A((size_t)M * K), B((size_t)N * K), Ah(A.size()), Al(A.size()), Bh(B.size()), Bl(B.size());
    srand(1);
    for (auto& v : A) v = (float)rand() / RAND_MAX * 2 - 1;
    for (auto& v : B) v = (float)rand() / RAND_MAX * 2 - 1;
    for (size_t i = 0; i < A.size(); ++i) { Ah[i] = tf32_rna(A[i]); Al[i] = tf32_rna(A[i] - Ah[i]); }
    for (size_t i = 0; i < B.size(); ++i) { Bh[i] = tf32_rna(B[i]); Bl[i] = tf32_rna(B[i] - Bh[i]); }
    float *dAh, *dAl, *dBh, *dBl, *dC;
    CK(cudaMalloc(&dAh, A.size() * 4)); CK(cudaMalloc(&dAl, A.size() * 4)); CK(cudaMalloc(&dBh, B.size() * 4)); CK(cudaMalloc(&dBl, B.size() * 4));
    CK(cudaMalloc(&dC, (size_t)M * N * 4));
    CK(cudaMemcpy(dAh, Ah.data(), A.size() * 4, cudaMemcpyHostToDevice)); CK(cudaMemcpy(dAl, Al.data(), A.size() * 4, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(dBh, Bh.data(), B.size() * 4, cudaMemcpyHostToDevice)); CK(cudaMemcpy(dBl, Bl.data(), B.size() * 4, cudaMemcpyHostToDevice));
    CK(cudaMemset(dC, 0xFF, (size_t)M * N * 4));
    CUtensorMap mAh = make_map(enc, dAh, M, K, BM), mAl = make_map(enc, dAl, M, K, BM), mBh = make_map(enc, dBh, N, K, BN), mBl = make_map(enc, dBl, N, K, BN);
    constexpr int smem = STAGES * (2 * BM * BK * 4 + 2 * BN * BK * 4) + 1024 + 256;
    CK(cudaFuncSetAttribute(k_tc<BN, STAGES>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    dim3 grid((M + BM - 1) / BM, (N + BN - 1) / BN);
    k_tc<BN, STAGES><<<grid, 192, smem>>>(mAh, mAl, mBh, mBl, dC, M, N, K);
    CK(cudaDeviceSynchronize());
    std::vector<float> C((size_t)M * N);
    CK(cudaMemcpy(C.data(), dC, C.size() * 4, cudaMemcpyDeviceToHost));
    double worst = 0, worst32 = 0, scale = 0;
    for (int i = 0; i < M; i += (M > 512 ? 37 : 1))
        for (int j = 0; j < N; ++j) {
            double ref = 0; float ref32 = 0;
            for (int k = 0; k < K; ++k) { ref += (double)A[(size_t)i * K + k] * B[(size_t)j * K + k]; ref32 += A[(size_t)i * K + k] * B[(size_t)j * K + k]; }
            worst = fmax(worst, fabs(C[(size_t)i * N + j] - ref)); worst32 = fmax(worst32, fabs(ref32 - ref)); scale = fmax(scale, fabs(ref));
        }
    printf("BN=%d M=%d N=%d K=%d: max abs err %.3e (fp32 sequential %.3e) scale %.3e -> rel %.3e\n", BN, M, N, K, worst, worst32, scale, worst / scale);
    // timing
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    for (int i = 0; i < 3; ++i) k_tc<BN, STAGES><<<grid, 192, smem>>>(mAh, mAl, mBh, mBl, dC, M, N, K);
    cudaEventRecord(e0);
    for (int i = 0; i < 10; ++i) k_tc<BN, STAGES><<<grid, 192, smem>>>(mAh, mAl, mBh, mBl, dC, M, N, K);
    cudaEventRecord(e1); CK(cudaDeviceSynchronize());
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    printf("   %.1f us per launch, %.1f TFLOP/s algorithmic\n", ms * 100, 2.0 * M * N * K / (ms / 10 * 1e-3) / 1e12);
    cudaEventRecord(e0);
    for (int i = 0; i < 10; ++i) k_tc<BN, STAGES><<<grid, 192, smem>>>(mAh, mAl, mBh, mBl, dC, M, N, K, 0);
    cudaEventRecord(e1); CK(cudaDeviceSynchronize());
    cudaEventElapsedTime(&ms, e0, e1);
    printf("   without epilogue stores: %.1f us per launch\n", ms * 100);
    cudaFree(dAh); cudaFree(dAl); cudaFree(dBh); cudaFree(dBl); cudaFree(dC);
}


template <int BN, int STAGES> static void run_wgrad(int M, int NO, int KI, int chunk) {
    EncodeFn enc = get_encode();
    std::vector<float> Y((size_t)M * NO), X((size_t)M * KI), Yh(Y.size()), Yl(Y.size()), Xh(X.size()), Xl(X.size());
    srand(2);
    for (auto& v : Y) v = (float)rand() / RAND_MAX * 2 - 1;
    for (auto& v : X) v = (float)rand() / RAND_MAX * 2 - 1;
    for (size_t i = 0; i < Y.size(); ++i) { Yh[i] = tf32_rna(Y[i]); Yl[i] = tf32_rna(Y[i] - Yh[i]); }
    for (size_t i = 0; i < X.size(); ++i) { Xh[i] = tf32_rna(X[i]); Xl[i] = tf32_rna(X[i] - Xh[i]); }
    float *dYh, *dYl, *dXh, *dXl, *dD;
    CK(cudaMalloc(&dYh, Y.size() * 4)); CK(cudaMalloc(&dYl, Y.size() * 4)); CK(cudaMalloc(&dXh, X.size() * 4)); CK(cudaMalloc(&dXl, X.size() * 4));
    CK(cudaMalloc(&dD, (size_t)NO * KI * 4));
    CK(cudaMemcpy(dYh, Yh.data(), Y.size() * 4, cudaMemcpyHostToDevice)); CK(cudaMemcpy(dYl, Yl.data(), Y.size() * 4, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(dXh, Xh.data(), X.size() * 4, cudaMemcpyHostToDevice)); CK(cudaMemcpy(dXl, Xl.data(), X.size() * 4, cudaMemcpyHostToDevice));
    CK(cudaMemset(dD, 0, (size_t)NO * KI * 4));
    CUtensorMap mYh = make_map(enc, dYh, M, NO, 32, CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B), mYl = make_map(enc, dYl, M, NO, 32, CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B), mXh = make_map(enc, dXh, M, KI, 32, CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B), mXl = make_map(enc, dXl, M, KI, 32, CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B);
    constexpr int smem = STAGES * (2 * BM * BK * 4 + 2 * BN * BK * 4) + 1024 + 256;
    CK(cudaFuncSetAttribute(k_tc_wgrad<BN, STAGES>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    dim3 grid((M + chunk - 1) / chunk, (NO + BM - 1) / BM, (KI + BN - 1) / BN);
    k_tc_wgrad<BN, STAGES><<<grid, 192, smem>>>(mYh, mYl, mXh, mXl, dD, M, NO, KI, chunk);
    CK(cudaDeviceSynchronize());
    std::vector<float> D((size_t)NO * KI);
    CK(cudaMemcpy(D.data(), dD, D.size() * 4, cudaMemcpyDeviceToHost));
    double worst = 0, scale = 0;
    for (int n = 0; n < NO; n += 7)
        for (int k = 0; k < KI; k += 5) {
            double ref = 0;
            for (int m = 0; m < M; ++m) ref += (double)Y[(size_t)m * NO + n] * X[(size_t)m * KI + k];
            worst = fmax(worst, fabs(D[(size_t)n * KI + k] - ref)); scale = fmax(scale, fabs(ref));
        }
    for (int t = 0; t < 4; ++t) {
        int n = t * 37 % NO, k = t * 53 % KI; double ref = 0;
        for (int m = 0; m < M; ++m) ref += (double)Y[(size_t)m * NO + n] * X[(size_t)m * KI + k];
        printf("   D[%d,%d] = %.5f ref %.5f | D[k,n]=%.5f\n", n, k, D[(size_t)n * KI + k], ref, (n < KI && k < NO) ? D[(size_t)k * KI + n] : 0.f);
    }
    printf("WGRAD BN=%d M=%d NO=%d KI=%d chunk=%d: max abs err %.3e scale %.3e -> rel %.3e\n", BN, M, NO, KI, chunk, worst, scale, worst / scale);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaEventRecord(e0);
    for (int i = 0; i < 10; ++i) k_tc_wgrad<BN, STAGES><<<grid, 192, smem>>>(mYh, mYl, mXh, mXl, dD, M, NO, KI, chunk);
    cudaEventRecord(e1); CK(cudaDeviceSynchronize());
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    printf("   %.1f us per launch, %.1f TFLOP/s algorithmic\n", ms * 100, 2.0 * M * NO * KI / (ms / 10 * 1e-3) / 1e12);
    cudaFree(dYh); cudaFree(dYl); cudaFree(dXh); cudaFree(dXl); cudaFree(dD);
}

int main() {
    run<128, 3>(256, 128, 64);
    run<128, 3>(98304, 256, 256);
    run<256, 2>(98304, 256, 256);
    run<128, 3>(98304, 128, 256);
    run<128, 3>(98304, 128, 128);
    return 0;
}
