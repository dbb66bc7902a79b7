// h2.cuh - the "h2" operand format of the learner GEMMs and the weight-gradient kernel built on it.
//
// h2 word: one fp32-class number s*x stored in the 4 bytes an fp32 would take, as TWO fp16 halves
//        bits [0,16)  hi = fp16(s x)                 (11 significant bits)
//        bits [16,32) lo = fp16(s x - hi)            (the next 11 bits)
// with a power-of-two scale s per tensor (activations 2^4, weights 2^10, gradients: a per-epoch device scalar chosen from a
// rigorous bound, see k_finalize_loss).  hi + lo carries ~22-24 significant bits wherever |s x| >= 2^-3 and an ABSOLUTE error
// <= 2^-25 below (fp16 subnormals) - i.e. at most 2^-24 of the tensor's largest magnitudes, the fp32 rounding class the
// reference's SGEMMs work in (utils/runner.py:132-133,148,163).
//
// Why: tcgen05.mma kind::f16 runs at TWICE the kind::tf32 rate and the two halves ride in ONE 32-bit word, so
//   * K-major operands (forward / input-gradient chains, mlp_chain_h2.cuh): a row of words IS a K-major fp16 row whose K index
//     alternates hi, lo.  A' * B'^T  = sum hi*hi + lo*lo   and   A' * B''^T = sum hi*lo + lo*hi  with B'' = B' with its halves
//     swapped: 2 MMAs per k-step instead of the 3 of the TF32 split, the word sits in one TMEM column (TS form) - no separate
//     lo operand, no shared-memory hand-over, no proxy fence.
//   * MN-major operands (weight gradients, below): the same words viewed as an fp16 matrix [rows, 2 cols] are an operand whose
//     M (or N) index alternates hi, lo: ONE MMA produces all four hi/lo products as a 2 x 2 block per output element, each in
//     its own TMEM cell (so the dominant hi*hi sum is never rounded together with the small terms), and the epilogue adds
//     the block.  No converter warps, no operand is split in shared memory (the TF32 kernel's measured limiter).
#pragma once
#include <cuda_fp16.h>

#include "gemm_tc.cuh"

namespace b200 {
namespace h2 {
using namespace tc;

static constexpr float S_ACT = 16.0f;        // hidden activations (ELU outputs >= -1; clamped at 65000 / 16)
static constexpr float S_W = 1024.0f;        // hidden-layer weights
static constexpr float S_X = 256.0f;         // network inputs (observations): with a scale of 1 the lo half of an input of magnitude 0.1 was an fp16 SUBNORMAL (absolute
                                             // precision 2^-24, i.e. 6e-7 of the value); |x| beyond 65000 / 256 = 254 is clamped
static constexpr float H2_MAX = 65000.0f;    // largest magnitude stored (fp16 max is 65504)

// device scalars of the format (workspace region `SC`): gradient scales of the two nets, chosen per epoch
enum { SC_SG_C = 0, SC_ISG_C = 1, SC_SG_A = 2, SC_ISG_A = 3, SC_COUNT = 16 };

__device__ __forceinline__ uint32_t pack(float x) {   // x = s * value, |x| <= H2_MAX
    // hi = x rounded to 11 significant bits (round half away, 2 ALU ops): exactly an fp16 in the normal range
    const float hi = __uint_as_float((__float_as_uint(x) + 0x1000u) & 0xFFFFE000u);
    const float lo = x - hi;
    uint32_t w;
    asm("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(w) : "f"(lo), "f"(hi));   // upper half <- lo, lower half <- hi
    return w;
}
__device__ __forceinline__ uint32_t pack_clamped(float x) { return pack(fminf(fmaxf(x, -H2_MAX), H2_MAX)); }
__device__ __forceinline__ uint32_t swap_halves(uint32_t w) { return __byte_perm(w, 0u, 0x1032); }
__device__ __forceinline__ float unpack(uint32_t w) {  // = s * value
    const float2 f = __half22float2(*reinterpret_cast<const __half2*>(&w));
    return f.x + f.y;
}

// instruction descriptor: D = f32, A = B = f16, M = 128
__host__ __device__ constexpr uint32_t idesc_f16(int n, bool mn_major) {
    return (1u << 4) | (mn_major ? ((1u << 15) | (1u << 16)) : 0u) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);
}
// MN-major 16-bit operand, SWIZZLE_128B: 64-element (128-byte) MN blocks `lbo` bytes apart, 8-row k groups 1024 B apart
__device__ __forceinline__ uint64_t desc_mn16(uint32_t saddr, uint32_t lbo) {
    return (uint64_t)((saddr >> 4) & 0x3FFF) | ((uint64_t)(lbo >> 4) << 16) | ((uint64_t)(1024 >> 4) << 32) | ((uint64_t)1 << 46) | ((uint64_t)2 << 61);
}
__device__ __forceinline__ void umma_f16(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(tmem_d), "l"(da), "l"(db), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void umma_f16_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t db, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}"
        ::"r"(tmem_d), "r"(tmem_a), "l"(db), "r"(idesc), "r"(accumulate) : "memory");
}

// bounded mbarrier wait: a protocol bug becomes a trapped kernel (a CUDA error at the next sync), never a hung GPU
__device__ unsigned int g_h2_error = 0;
__device__ __forceinline__ void mbar_wait_wd(uint64_t* b, uint32_t parity, int code) {
    uint32_t done;
    const uint32_t a = smem_u32(b);
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(done) : "r"(a), "r"(parity) : "memory");
    if (done) return;
    const long long t0 = clock64();
    do {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(done) : "r"(a), "r"(parity) : "memory");
        if (!done && clock64() - t0 > 4000000000ll) {
            atomicExch(&g_h2_error, (unsigned)code);
            printf("h2 watchdog: block %d thread %d barrier code %d parity %u\n", (int)blockIdx.x, (int)threadIdx.x, code, parity);
            __trap();
        }
    } while (!done);
}

// =====================================================================================================================
// weight gradients: D[n, k] = sum over rows m of dY[m, n] X[m, k]   for ALL SIX hidden-layer matrices of an epoch in ONE launch
// =====================================================================================================================
// Work: every matrix is cut into 128 x 128 (or 128 x 64) output tiles, every tile's rows into `parts` ranges so that the
// whole launch is one wave of <= SM-count CTAs of (nearly) equal cost; each CTA stores its partial tile, k_wgrad_h2_reduce sums the
// parts, undoes the scales and adds into the gradient.  Per CTA:
//   warp 0   TMA producer: a stage = 32 rows of the dY words [32 x 128] and of the X words [32 x 128 | 64] (3-D tensor maps
//            (32 words, rows, 32-word column chunks), 128-byte swizzle: chunk c lands at c * 4 KB, row r at r * 128 B)
//   warp 1   MMA issuer: per 16 rows two tcgen05.mma kind::f16, M = 128 (= 64 real output rows x {hi, lo}), N = 256 | 128
//            (= 128 | 64 real columns x {hi, lo}), K = 16, MN-major operands straight from the landed words -> TMEM [0,256) and
//            [256,512)
//   warps 2-17  flush: every FLUSH stages (TMEM accumulation truncates on every add; short chains + round-to-nearest fp32
//            adds in registers keep the 98k-row sums at fp32 accuracy) the accumulators are drained into registers: a thread
//            owns one TMEM lane and 128 columns, adds each (hi-col, lo-col) pair -> 64 running sums; at the end lane pairs
//            (hi row, lo row) are combined with one shuffle and the 128 x 128 partial tile is stored.
struct WgTile {
    int job;          // which (dY, X) pair
    int n0, k0;       // first real output row / column of the tile
    int first;        // first CTA of the tile
    int parts;        // CTAs of the tile
    int chunk;        // rows per CTA (multiple of 32)
};
static constexpr int WG_MAX_TILES = 16;
struct alignas(64) WgParams {
    CUtensorMap mY[6], mX[6];
    WgTile tile[WG_MAX_TILES];
    int ntiles, M;
    int kchunks[6];     // 32-word column chunks of the X tile: 4 (128 columns) or 2 (64)
    float* P;           // partial tiles [CTA][128][128] fp32
};
static constexpr int WG2_STAGES = 6, WG2_ROWS = 32, WG2_BOX = WG2_ROWS * 128, WG2_A = 4 * WG2_BOX, WG2_STAGE = 2 * WG2_A;
static constexpr int WG2_FLUSH = 32;       // stages per TMEM accumulation chain (64 MMAs per accumulator)
static constexpr int WG2_EPI0 = 2, WG2_THREADS = 32 * (WG2_EPI0 + 16);
static constexpr int WG2_SMEM = WG2_STAGES * WG2_STAGE + 256 + 1024;

__global__ void __launch_bounds__(WG2_THREADS, 1) k_wgrad_h2(const __grid_constant__ WgParams P) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    uint64_t* full = (uint64_t*)(smem + WG2_STAGES * WG2_STAGE);
    uint64_t* empty = full + WG2_STAGES;
    uint64_t* tfull = empty + WG2_STAGES;     // accumulators complete (one flush interval)
    uint64_t* tempty = tfull + 1;             // accumulators drained
    uint32_t* tmem_slot = (uint32_t*)(tempty + 1);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t sbase = smem_u32(smem);
    // this CTA's tile and row range
    int ti = 0;
    for (int t = 1; t < P.ntiles; ++t) if ((int)blockIdx.x >= P.tile[t].first) ti = t;
    const WgTile& T = P.tile[ti];
    const int part = (int)blockIdx.x - T.first;
    const int r_begin = part * T.chunk, r_end = min(P.M, r_begin + T.chunk);
    const int nk = (part < T.parts && r_end > r_begin) ? (r_end - r_begin + WG2_ROWS - 1) / WG2_ROWS : 0;
    const int kch = P.kchunks[T.job];
    const int nflush = (nk + WG2_FLUSH - 1) / WG2_FLUSH;

    if (warp == 0 && lane == 0) {
        for (int s = 0; s < WG2_STAGES; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
        mbar_init(tfull, 1); mbar_init(tempty, 16);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(512u));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    asm volatile("tcgen05.fence::before_thread_sync;");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;");
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        if (lane == 0) {
            const uint32_t bytes = (uint32_t)(WG2_A + kch * WG2_BOX);
            for (int kb = 0; kb < nk; ++kb) {
                const uint32_t s = kb % WG2_STAGES, ph = (kb / WG2_STAGES) & 1;
                mbar_wait_wd(&empty[s], ph ^ 1, 900 + (int)s);
                const uint32_t st = sbase + s * WG2_STAGE;
                mbar_expect_tx(&full[s], bytes);
                const int r = r_begin + kb * WG2_ROWS;
                tma_load_3d(&P.mY[T.job], &full[s], st, 0, r, T.n0 / 32);
                tma_load_3d(&P.mX[T.job], &full[s], st + WG2_A, 0, r, T.k0 / 32);
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            const uint32_t idesc = idesc_f16(kch * 64, true);
            for (int kb = 0; kb < nk; ++kb) {
                const uint32_t s = kb % WG2_STAGES, ph = (kb / WG2_STAGES) & 1;
                const int fi = kb / WG2_FLUSH;
                const bool first = (kb % WG2_FLUSH) == 0;
                if (first && fi > 0) {
                    umma_commit(tfull);                                   // interval fi - 1 complete ...
                    mbar_wait_wd(tempty, (uint32_t)((fi - 1) & 1), 910);  // ... and drained into the flush warps' registers
                    asm volatile("tcgen05.fence::after_thread_sync;");
                }
                mbar_wait_wd(&full[s], ph, 920 + (int)s);
                asm volatile("tcgen05.fence::after_thread_sync;");
                const uint32_t a = sbase + s * WG2_STAGE, b = a + WG2_A;
                const uint64_t da0 = desc_mn16(a, WG2_BOX), db0 = desc_mn16(b, WG2_BOX);
#pragma unroll
                for (int k = 0; k < WG2_ROWS / 16; ++k) {
                    const uint64_t koff = (uint64_t)(k * 2048 >> 4);   // 16 rows (samples) = two 8-row swizzle groups; start-address field counts 16 bytes
                    const uint32_t acc = (first && k == 0) ? 0u : 1u;
                    umma_f16(tmem_base, da0 + koff, db0 + koff, idesc, acc);
                    umma_f16(tmem_base + 256, da0 + (2 * WG2_BOX >> 4) + koff, db0 + koff, idesc, acc);
                }
                umma_commit(&empty[s]);
            }
            if (nk > 0) umma_commit(tfull);
        }
    } else {
        // ===== flush warps: TMEM lane quarter q, 128-column group cg: accumulator cg / 2, real columns (cg & 1) * 64 .. + 63 =====
        const int q = warp & 3, cg = (warp - WG2_EPI0) >> 2;
        const bool active = (cg & 1) * 128 < kch * 64;           // (64-column tiles use only the first 128 columns of each accumulator)
        const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(cg * 128);
        float acc[64];
#pragma unroll
        for (int j = 0; j < 64; ++j) acc[j] = 0.0f;
        for (int fi = 0; fi < nflush; ++fi) {
            mbar_wait_wd(tfull, (uint32_t)(fi & 1), 930);
            asm volatile("tcgen05.fence::after_thread_sync;");
            if (active) {
#pragma unroll
                for (int c = 0; c < 4; ++c) {
                    uint32_t r[32];
                    tmem_ld32(taddr + c * 32, r);
#pragma unroll
                    for (int j = 0; j < 16; ++j) acc[c * 16 + j] += __uint_as_float(r[2 * j]) + __uint_as_float(r[2 * j + 1]);
                }
            }
            asm volatile("tcgen05.fence::before_thread_sync;");
            __syncwarp();
            if (lane == 0) mbar_arrive(tempty);
        }
        if (active && part < T.parts) {
            // lanes (2i, 2i + 1) hold the hi-row and lo-row parts of real output row i: one shuffle, then each lane stores half of the row's
            // 64 columns.  Partial tile [CTA][128][128]: real row (cg / 2) * 64 + (q * 32 + lane) / 2, columns (cg & 1) * 64 ...
            float* dst = P.P + ((size_t)blockIdx.x * 128 + (cg >> 1) * 64 + ((q * 32 + lane) >> 1)) * 128 + (cg & 1) * 64 + (lane & 1) * 32;
#pragma unroll
            for (int j = 0; j < 64; ++j) acc[j] += __shfl_xor_sync(0xffffffffu, acc[j], 1);
            float o[32];
#pragma unroll
            for (int j = 0; j < 32; ++j) o[j] = (lane & 1) ? acc[32 + j] : acc[j];
#pragma unroll
            for (int j = 0; j < 32; j += 8) stg_v8(dst + j, o + j);
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;");
    __syncthreads();
    if (warp == 1) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u));
}

// sum of the partial tiles -> gradient: dW[n, k] += (sum over parts) / (s_dY * s_X); one thread per output element, all six matrices
struct WgRedJob {
    float* D;            // [n_out, ldd] gradient
    const float* isg;    // device: 1 / (gradient scale of this net)
    float inv_sx;        // 1 / (scale of the X operand)
    int n_out, k_valid, ldd;
    int tile0;           // first tile of this job in WgParams::tile (tiles ordered n-major, then k)
    int tiles_k;         // tiles along k
    int first;           // first linear output element of this job
};
struct WgRedJobs {
    WgRedJob job[6];
    WgTile tile[WG_MAX_TILES];
    const float* P;
    int count, total;
};
__global__ void __launch_bounds__(256) k_wgrad_h2_reduce(const WgRedJobs J) {
    const int idx = blockIdx.x * 256 + threadIdx.x;
    if (idx >= J.total) return;
    int j = 0;
#pragma unroll
    for (int t = 1; t < 6; ++t)
        if (t < J.count && idx >= J.job[t].first) j = t;
    const WgRedJob& w = J.job[j];
    const int e = idx - w.first;
    const int row = e / w.k_valid, col = e - row * w.k_valid;
    const WgTile& T = J.tile[w.tile0 + (row >> 7) * w.tiles_k + (col >> 7)];
    const float* src = J.P + ((size_t)T.first * 128 + (row & 127)) * 128 + (col & 127);
    float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
    int p = 0;
    for (; p + 4 <= T.parts; p += 4) {
        s0 += src[(size_t)p * 16384]; s1 += src[(size_t)(p + 1) * 16384]; s2 += src[(size_t)(p + 2) * 16384]; s3 += src[(size_t)(p + 3) * 16384];
    }
    for (; p < T.parts; ++p) s0 += src[(size_t)p * 16384];
    w.D[(size_t)row * w.ldd + col] += ((s0 + s1) + (s2 + s3)) * (w.isg[0] * w.inv_sx);
}

}  // namespace h2
}  // namespace b200
