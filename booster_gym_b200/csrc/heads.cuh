// heads.cuh - backward of the output heads (critic 128 -> 1, actor 128 -> 12; utils/model.py:15,25) as ONE streaming pass over
// the last hidden activations, fed by bulk asynchronous copies.
//
//   dz3[m,k] = (sum_j D[m,j] W[j,k]) ELU'(H[m,k])      -> h2 words of sg * dz3 (operand of the backward chain and of k_wgrad_h2)
//   dW[j,k] += sum_m D[m,j] H[m,k]     db[j] += sum_m D[m,j]     db3[k] += sum_m dz3[m,k]   (bias gradient of hidden layer 3)
//
// The first version (k_actor_head_bwd / k_value_head_bwd, learner_kernels.cu) loads each row straight from HBM inside the
// arithmetic loop: 16 warps per SM x 2 rows in flight = 16 KB per SM cannot cover the ~1 us HBM latency (measured 69 + 24 us for
// 110 MB that take 17 us at the HBM peak).  Here a producer warp streams [64 rows x 512 B] tiles (contiguous in HBM: rows are
// dense) into a two-stage shared-memory ring with cp.async.bulk + mbarrier (two blocks per SM: up to 140 KB in flight per SM) and
// seven consumer warps (256 threads with the producer: 128 registers at two blocks per SM) run the same arithmetic on shared memory.
#pragma once
#include "h2.cuh"

namespace b200 {
namespace heads {
using namespace tc;

static constexpr int HB2_ROWS = 64, HB2_STAGES = 2, HB2_WARPS = 7, HB2_THREADS = 32 * (HB2_WARPS + 1);

__device__ __forceinline__ void bulk_load(uint32_t dst, const void* src, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(dst), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}

template <int J> struct HeadSmem {
    static constexpr int H_BYTES = HB2_ROWS * 512, D_BYTES = HB2_ROWS * J * 4, STAGE = H_BYTES + ((D_BYTES + 127) / 128) * 128;
    static constexpr int W_OFF = HB2_STAGES * STAGE, BAR_OFF = W_OFF + J * 512, TOTAL = BAR_OFF + 64;
};

template <int J>
__global__ void __launch_bounds__(HB2_THREADS, 2) k_head_bwd_pipe(const float* __restrict__ H, const float* __restrict__ W, const float* __restrict__ D,
                                                                  int n, uint32_t* __restrict__ DZ, float* __restrict__ dW, float* __restrict__ db,
                                                                  float* __restrict__ db_prev, const float* __restrict__ sg) {
    using S = HeadSmem<J>;
    extern __shared__ __align__(128) uint8_t smem[];
    uint64_t* full = (uint64_t*)(smem + S::BAR_OFF);
    uint64_t* empty = full + HB2_STAGES;
    float4* sw = (float4*)(smem + S::W_OFF);   // [J][32]
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t sbase = smem_u32(smem);
    const int tiles = (n + HB2_ROWS - 1) / HB2_ROWS;
    for (int i = threadIdx.x; i < J * 32; i += HB2_THREADS) sw[i] = reinterpret_cast<const float4*>(W)[i];
    if (threadIdx.x == 0) {
        for (int s = 0; s < HB2_STAGES; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], HB2_WARPS); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    float4 acc[J];
#pragma unroll
    for (int j = 0; j < J; ++j) acc[j] = make_float4(0.f, 0.f, 0.f, 0.f);
    float4 ap = make_float4(0.f, 0.f, 0.f, 0.f);
    float ab = 0.0f;   // lane j < J sums D[:, j]
    if (warp == HB2_WARPS) {
        // ===== producer: bulk copies of the H tile and the D tile (rows are dense: one contiguous range each) =====
        if (lane == 0) {
            uint32_t it = 0;
            for (int t = blockIdx.x; t < tiles; t += gridDim.x, ++it) {
                const uint32_t s = it % HB2_STAGES, ph = (it / HB2_STAGES) & 1;
                h2::mbar_wait_wd(&empty[s], ph ^ 1, 950);
                const int r0 = t * HB2_ROWS, rows = min(HB2_ROWS, n - r0);
                const uint32_t hb = (uint32_t)rows * 512u, dbytes = ((uint32_t)(rows * J * 4) + 15u) & ~15u;   // (the workspace pads D beyond n)
                mbar_expect_tx(&full[s], hb + dbytes);
                bulk_load(sbase + s * S::STAGE, H + (size_t)r0 * 128, hb, &full[s]);
                bulk_load(sbase + s * S::STAGE + S::H_BYTES, D + (size_t)r0 * J, dbytes, &full[s]);
            }
        }
    } else {
        const float gs = sg[0];
        // this lane's 4 columns of every head row stay in REGISTERS: re-reading them from shared memory for every row (J LDS.128 per
        // row and warp) made the J = 12 kernels shared-memory-bandwidth bound (52 clk of the LSU pipe per row: 18 us of 37 measured)
        float4 wr[J];
#pragma unroll
        for (int j = 0; j < J; ++j) wr[j] = sw[j * 32 + lane];
        uint32_t it = 0;
        for (int t = blockIdx.x; t < tiles; t += gridDim.x, ++it) {
            const uint32_t s = it % HB2_STAGES, ph = (it / HB2_STAGES) & 1;
            h2::mbar_wait_wd(&full[s], ph, 951);
            const int r0 = t * HB2_ROWS, rows = min(HB2_ROWS, n - r0);
            const float4* hs = reinterpret_cast<const float4*>(smem + s * S::STAGE);
            const float* ds = reinterpret_cast<const float*>(smem + s * S::STAGE + S::H_BYTES);
#pragma unroll 1
            for (int r = warp; r < rows; r += HB2_WARPS) {
                const float4 h = hs[r * 32 + lane];
                float4 g = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
                for (int j = 0; j < J; ++j) {
                    const float4 wj = wr[j];
                    const float dj = ds[r * J + j];   // (broadcast load, consumed at once: 2 x J float4 of weights / sums live in registers)
                    g.x = fmaf(dj, wj.x, g.x); g.y = fmaf(dj, wj.y, g.y); g.z = fmaf(dj, wj.z, g.z); g.w = fmaf(dj, wj.w, g.w);
                    acc[j].x = fmaf(dj, h.x, acc[j].x); acc[j].y = fmaf(dj, h.y, acc[j].y);
                    acc[j].z = fmaf(dj, h.z, acc[j].z); acc[j].w = fmaf(dj, h.w, acc[j].w);
                }
                float4 dh;
                dh.x = g.x * ((h.x > 0.0f) ? 1.0f : (h.x + 1.0f));
                dh.y = g.y * ((h.y > 0.0f) ? 1.0f : (h.y + 1.0f));
                dh.z = g.z * ((h.z > 0.0f) ? 1.0f : (h.z + 1.0f));
                dh.w = g.w * ((h.w > 0.0f) ? 1.0f : (h.w + 1.0f));
                uint4 wd;
                wd.x = h2::pack(dh.x * gs); wd.y = h2::pack(dh.y * gs); wd.z = h2::pack(dh.z * gs); wd.w = h2::pack(dh.w * gs);
                reinterpret_cast<uint4*>(DZ + (size_t)(r0 + r) * 128)[lane] = wd;
                ap.x += dh.x; ap.y += dh.y; ap.z += dh.z; ap.w += dh.w;
                if (lane < J) ab += ds[r * J + lane];
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(&empty[s]);
        }
    }
    // ---- block combine: (J dW rows + the db3 column sums) x 128 columns, 8 warp partials each, in passes of 7 rows through the (now idle) ring
    __syncthreads();
    float4* red = reinterpret_cast<float4*>(smem);          // [HB2_WARPS][7][32]
    float* redb = reinterpret_cast<float*>(smem + HB2_WARPS * 7 * 512);   // [HB2_WARPS][16]
    if (warp < HB2_WARPS && lane < J) redb[warp * 16 + lane] = ab;
    constexpr int NOUT = J + 1;
#pragma unroll
    for (int j0 = 0; j0 < NOUT; j0 += 7) {
        const int cnt = (NOUT - j0) < 7 ? (NOUT - j0) : 7;
        if (j0) __syncthreads();
        if (warp < HB2_WARPS) {
#pragma unroll
            for (int j = 0; j < 7; ++j) {
                if (j < cnt) {
                    float4 v = ap;
#pragma unroll
                    for (int jj = 0; jj < J; ++jj) if (jj == j0 + j) v = acc[jj];
                    red[(warp * 7 + j) * 32 + lane] = v;
                }
            }
        }
        __syncthreads();
        for (int o = threadIdx.x; o < cnt * 128; o += HB2_THREADS) {
            const int j = o >> 7, k = o & 127;
            double sacc = 0.0;
#pragma unroll
            for (int v = 0; v < HB2_WARPS; ++v) sacc += (double)reinterpret_cast<const float*>(&red[(v * 7 + j) * 32])[k];
            if (j0 + j < J) atomicAdd(dW + (j0 + j) * 128 + k, (float)sacc);
            else atomicAdd(db_prev + k, (float)sacc);
        }
    }
    if (threadIdx.x < J) {
        double sb = 0.0;
        for (int v = 0; v < HB2_WARPS; ++v) sb += (double)redb[v * 16 + threadIdx.x];
        atomicAdd(db + threadIdx.x, (float)sb);
    }
}

// forward of the heads with the same feed: OUT[m, j] = b[j] + sum_k H[m,k] W[j,k]  (V for J = 1, mu for J = 12).  One warp per row, a
// float4 of hidden units per lane, the J partial dot products combined by a 16-value butterfly reduce-scatter (lane l ends up with output l).
template <int J>
__global__ void __launch_bounds__(HB2_THREADS, 2) k_head_fwd_pipe(const float* __restrict__ H, const float* __restrict__ W, const float* __restrict__ b,
                                                                  int n, float* __restrict__ OUT) {
    constexpr int STAGE = HB2_ROWS * 512, W_OFF = HB2_STAGES * STAGE, BAR_OFF = W_OFF + J * 512;
    extern __shared__ __align__(128) uint8_t smem[];
    uint64_t* full = (uint64_t*)(smem + BAR_OFF);
    uint64_t* empty = full + HB2_STAGES;
    float4* sw = (float4*)(smem + W_OFF);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t sbase = smem_u32(smem);
    const int tiles = (n + HB2_ROWS - 1) / HB2_ROWS;
    for (int i = threadIdx.x; i < J * 32; i += HB2_THREADS) sw[i] = reinterpret_cast<const float4*>(W)[i];
    if (threadIdx.x == 0) {
        for (int s = 0; s < HB2_STAGES; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], HB2_WARPS); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (warp == HB2_WARPS) {
        if (lane == 0) {
            uint32_t it = 0;
            for (int t = blockIdx.x; t < tiles; t += gridDim.x, ++it) {
                const uint32_t s = it % HB2_STAGES, ph = (it / HB2_STAGES) & 1;
                h2::mbar_wait_wd(&empty[s], ph ^ 1, 960);
                const int r0 = t * HB2_ROWS, rows = min(HB2_ROWS, n - r0);
                mbar_expect_tx(&full[s], (uint32_t)rows * 512u);
                bulk_load(sbase + s * STAGE, H + (size_t)r0 * 128, (uint32_t)rows * 512u, &full[s]);
            }
        }
        return;
    }
    const float bj = lane < J ? b[lane] : 0.0f;
    float4 wr[J];   // this lane's 4 columns of every head row, in registers (see k_head_bwd_pipe)
#pragma unroll
    for (int j = 0; j < J; ++j) wr[j] = sw[j * 32 + lane];
    uint32_t it = 0;
    for (int t = blockIdx.x; t < tiles; t += gridDim.x, ++it) {
        const uint32_t s = it % HB2_STAGES, ph = (it / HB2_STAGES) & 1;
        h2::mbar_wait_wd(&full[s], ph, 961);
        const int r0 = t * HB2_ROWS, rows = min(HB2_ROWS, n - r0);
        const float4* hs = reinterpret_cast<const float4*>(smem + s * STAGE);
#pragma unroll 2
        for (int r = warp; r < rows; r += HB2_WARPS) {
            const float4 h = hs[r * 32 + lane];
            if (J == 1) {
                const float4 w4 = wr[0];
                float p = fmaf(h.x, w4.x, fmaf(h.y, w4.y, fmaf(h.z, w4.z, h.w * w4.w)));
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) p += __shfl_xor_sync(0xffffffffu, p, o);
                if (lane == 0) OUT[r0 + r] = p + bj;
            } else {
                float p[16];
#pragma unroll
                for (int j = 0; j < 16; ++j) {
                    if (j < J) {
                        const float4 w4 = wr[j < J ? j : 0];
                        p[j] = fmaf(h.x, w4.x, fmaf(h.y, w4.y, fmaf(h.z, w4.z, h.w * w4.w)));
                    } else {
                        p[j] = 0.0f;
                    }
                }
#pragma unroll
                for (int half = 8; half >= 1; half >>= 1) {
                    const bool upper = (lane & half) != 0;
#pragma unroll
                    for (int j = 0; j < half; ++j) {
                        const float send = upper ? p[j] : p[j + half];
                        const float keep = upper ? p[j + half] : p[j];
                        p[j] = keep + __shfl_xor_sync(0xffffffffu, send, half);
                    }
                }
                const float tot = p[0] + __shfl_xor_sync(0xffffffffu, p[0], 16);   // lane l (mod 16) holds output l
                if (lane < J) OUT[(size_t)(r0 + r) * J + lane] = tot + bj;
            }
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(&empty[s]);
    }
}

}  // namespace heads
}  // namespace b200
