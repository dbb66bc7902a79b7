// mlp_chain_h2.cuh - the fused layer chains of mlp_chain.cuh on the h2 operand format (h2.cuh): tcgen05.mma kind::f16.
//
//   forward   X[128,64] -> h1 = ELU(X W1^T + b1) -> h2 = ELU(h1 W2^T + b2) -> h3 = ELU(h2 W3^T + b3)
//   backward  dz3[128,128] -> dz2 = (dz3 W3) * ELU'(h2) -> dz1 = (dz2 W2) * ELU'(h1)        (+ bias gradients = column sums)
//
// Every operand element is ONE 32-bit word {fp16 hi, fp16 lo} (scaled by a power of two).  A row of words is a K-major fp16 row
// whose K index alternates hi, lo, so with the weights staged twice - B' = words, B'' = words with the halves swapped -
//        D_main  += A' B'^T   = sum_k  a_hi b_hi + a_lo b_lo
//        D_small += A' B''^T  = sum_k  a_hi b_lo + a_lo b_hi
// is the full fp32-class product in TWO MMAs per k-step at twice the TF32 rate (the 3xTF32 chain needed three at the TF32 rate).
// The hand-over between layers is one tcgen05.st: the epilogue thread that owns (row, column c) of layer l's accumulator writes
// the packed word of its activation back to column c - the TMEM cell a kind::f16 TS-form MMA reads as the K pair (2c, 2c+1) of its
// A operand.  No lo operand, no shared-memory ring, no proxy fence: what is left of a hand-over is tcgen05.ld -> bias, ELU, pack ->
// tcgen05.st -> tcgen05.wait::st -> mbarrier arrive.
// TMEM maps are those of mlp_chain.cuh.  The activations go to HBM as the same words (4 B per element, what fp32 took): the
// backward chain reads them back as such, and the weight-gradient kernel takes them as MN-major operands without any conversion.
#pragma once
#include "h2.cuh"
#include "mlp_chain.cuh"

namespace b200 {
namespace chain2 {
using namespace tc;
using chain::TileSeq;
using chain::tmem_ld8;
using chain::tmem_ld8_nowait;
using chain::tmem_st8;
using chain::tmem_wait_ld16;
using chain::tmem_wait_ld8;
using chain::tmem_wait_st;
using h2::mbar_wait_wd;
using h2::umma_f16;
using h2::umma_f16_ts;
#ifdef B200_CHAIN_TL
using chain::g_chain_tl;
#endif

static constexpr int UNIT_BYTES = 32768;          // one weight-ring unit: [256 x 32] words (B' OR B'') or [128 x 32] B' + B''
static constexpr int KB_BYTES = BM * BK * 4;      // one [128 x 32] word k-block tile
static constexpr int HAND = 8;                    // hand-over barriers (>= k-blocks of a layer: the epilogue cannot run further ahead)
static constexpr int EPI_W = 16;

__host__ __device__ constexpr uint32_t idesc_f16_k(int n) {   // K-major operands, D = f32, A = B = f16, M = 128
    return (1u << 4) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);
}

// ELU in 5 straight-line instructions (see mlp_chain.cuh: absolute error <= ~3e-7)
__device__ __forceinline__ float elu_fast5(float x) {
    float e;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(x * 1.4426950408889634f));
    return (x > 0.0f) ? x : (e - 1.0f);
}

// the hand-over: this warp's 32 rows x 8 columns of the next layer's A operand are in TMEM
__device__ __forceinline__ void handover(uint64_t* bar, int lane) {
    tmem_wait_st();
    asm volatile("tcgen05.fence::before_thread_sync;");
    __syncwarp();
    if (lane == 0) mbar_arrive(bar);
}

// Results leave the SM through TMA: a thread owns one ROW of the tile (its TMEM lane), so a direct global store touches 32 different
// 128-byte lines per warp instruction - measured 68 % busy LSU data pipe, with the bias loads queueing behind the stores.  Instead
// the 16 warps write their 32-byte row pieces into a [128 x 32]-word staging tile in shared memory (128-byte swizzle, conflict-free
// 16-byte stores) and one thread of a dedicated warp sends the 16 KB tile out as ONE bulk tensor store (rows beyond the tensor are clipped).
#ifndef H2_O_STAGES
#define H2_O_STAGES 2
#endif
static constexpr int O_STAGES = H2_O_STAGES;
// MC = 1: clusters of two CTAs share the WEIGHT stream.  Every CTA re-reads all weights of a net for each of its row tiles (0.9 MB
// per critic tile): measured, with MMAs, arithmetic and stores switched off the forward kernel still took 107 of its 190 us - the
// L2 -> SM fill (~42 B/clk/SM).  The two CTAs of a cluster walk the same tile sequence on neighbouring row tiles; each loads HALF of
// every weight k-block and multicasts it into both shared memories (cp.async.bulk.tensor ... .multicast::cluster), a ring stage is
// free once BOTH tensor cores have read it (tcgen05.commit ... .multicast::cluster on the stage's `empty` barrier).  Everything else
// (input / aux tiles, TMEM, epilogue, stores) stays private to the CTA; the MMAs stay cta_group::1.
__device__ __forceinline__ void tma_load_2d_mc(const CUtensorMap* map, uint64_t* bar, uint32_t dst, int c0, int c1) {
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster [%0], [%1, {%3, %4}], [%2], %5;"
                 ::"r"(dst), "l"((uint64_t)map), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "h"((uint16_t)3) : "memory");
}
__device__ __forceinline__ void umma_commit_mc(uint64_t* bar) {   // arrives on `bar` in BOTH CTAs of the cluster
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
                 ::"r"(smem_u32(bar)), "h"((uint16_t)3) : "memory");
}
template <int MC> __device__ __forceinline__ void release_unit(uint64_t* bar) {
    if (MC) umma_commit_mc(bar); else umma_commit(bar);
}
// one weight-ring unit: [rows x 32] words of map `m` at k-block kb -> dst; MC: this CTA loads its half of the rows for both CTAs
template <int MC> __device__ __forceinline__ void load_rows(const CUtensorMap* m, uint64_t* bar, uint32_t dst, int kb, int rows, uint32_t rank) {
    if (MC) tma_load_2d_mc(m, bar, dst + rank * (uint32_t)(rows / 2) * 128u, kb * BK, (int)rank * (rows / 2));
    else tma_load_2d(m, bar, dst, kb * BK, 0);
}
#ifdef H2_HALF_W   // measurement build (results are garbage): only the B' half of every weight k-block is fetched - what a CTA pair's MMA would stream per SM
#define H2_SKIP_BPP 1
#else
#define H2_SKIP_BPP 0
#endif
__device__ __forceinline__ void tma_store_2d(const CUtensorMap* map, uint32_t src, int c0, int c1) {
    asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];"
                 ::"l"((uint64_t)map), "r"(src), "r"(c0), "r"(c1) : "memory");
    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
}
// this thread's 8 words (columns c8 * 8 ...) of row `rit` -> staging tile (16-byte chunk index XOR row % 8)
__device__ __forceinline__ void stage_words(uint32_t tile, int rit, int c8, const uint32_t* w) {
    const uint32_t rbase = tile + (uint32_t)rit * 128u, sw = (uint32_t)(rit & 7);
    sts_v4(rbase + ((((uint32_t)(2 * c8)) ^ sw) << 4), make_float4(__uint_as_float(w[0]), __uint_as_float(w[1]), __uint_as_float(w[2]), __uint_as_float(w[3])));
    sts_v4(rbase + ((((uint32_t)(2 * c8 + 1)) ^ sw) << 4), make_float4(__uint_as_float(w[4]), __uint_as_float(w[5]), __uint_as_float(w[6]), __uint_as_float(w[7])));
}
// the staging tile is complete for this warp: make the writes visible to the async proxy, then count the warp in
__device__ __forceinline__ void stage_done(uint64_t* bar, int lane) {
    fence_async_smem();
    __syncwarp();
    if (lane == 0) mbar_arrive(bar);
}

// =====================================================================================================================
// forward
// =====================================================================================================================
struct alignas(64) FwdNet {
    CUtensorMap mX;                       // input words [rows, 64], box 32 x 128
    CUtensorMap mW1a, mW1b;               // [256, 64]   box 32 x 256     (a = B', b = B'': halves swapped)
    CUtensorMap mW2a, mW2b;               // [n2, 256]   box 32 x n2
    CUtensorMap mW3a, mW3b;               // [128, n2]   box 32 x 128
    CUtensorMap mH1, mH2, mH3;            // outputs: h1 words [rows,256], h2 words [rows,n2], h3 fp32 [rows,128]; box 32 x 128
    const float *b1s, *b2s, *b3;          // biases of layers 1 and 2 pre-scaled by S_ACT (k_weight_prep), layer 3 as it is
    int rows, n2, pad0_, pad_;
    // Midpoint compensation of the truncating accumulator.  A tcgen05.mma adds into TMEM with truncation toward zero, so an accumulator
    // built by n accumulating MMAs sits between 0 and n ulp BELOW its exact magnitude - a bias that, unlike round-to-nearest errors, does
    // not average out over the batch (measured: mu -7.3e-7 of its mean magnitude, and the actor's gradients inherit it amplified by 1 /
    // sigma).  gain[l] = 1 + n_l 2^-26 (n_l = accumulating MMAs of layer l with the dominant terms present) moves the result to the middle
    // of that interval: mean signed error of mu +5e-8, max error halved (h2::midpoint_gain, set by the host; B200_H2_MIDPOINT=0: all 1).
    float gain[4];
};
struct alignas(64) FwdParams { FwdNet net[2]; };

static constexpr int F_UNITS = O_STAGES == 2 ? 5 : 4;
static constexpr int F_EPI0 = 3;                           // first epilogue warp
static constexpr int F_THREADS = 32 * (F_EPI0 + EPI_W);    // TMA loads, MMA, TMA stores, 16 epilogue warps
static constexpr int F_X = 0, F_B = 2 * KB_BYTES, F_OUT = F_B + F_UNITS * UNIT_BYTES, F_BAR = F_OUT + O_STAGES * KB_BYTES,
                     F_SMEM = F_BAR + 256 + 1024;
static_assert(F_SMEM <= 232448, "shared memory budget");

// G = epilogue warp groups: group j owns the k-blocks kb = j (mod G) of every layer and walks a k-block's 32 columns in G chunks of
// 8 per thread, so G hand-overs are in flight at once and one group's fixed latencies (tcgen05.ld, tcgen05.wait::st, the mbarrier
// round trip to the MMA thread) hide behind the other groups' arithmetic.  G = 1: all 16 warps on the same k-block.
template <int G, int MC>
__global__ void __launch_bounds__(F_THREADS, 1) k_mlp_fwd_h2(const __grid_constant__ FwdParams P) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    uint64_t* x_full = (uint64_t*)(smem + F_BAR);
    uint64_t* x_empty = x_full + 1;
    uint64_t* b_full = x_empty + 1;          // [F_UNITS]
    uint64_t* b_empty = b_full + F_UNITS;    // [F_UNITS]
    uint64_t* a_full = b_empty + F_UNITS;    // [HAND] a k-block of the next layer's A operand is in TMEM
    uint64_t* accf = a_full + HAND;          // [3] layer l's accumulator complete
    uint64_t* drained = accf + 3;            // every epilogue warp has read the tile's last accumulator (its columns are layer 2's of the next tile)
    uint64_t* o_full = drained + 1;          // [O_STAGES] an output staging tile is complete
    uint64_t* o_empty = o_full + O_STAGES;   // [O_STAGES] ... and has been read by its bulk store
    uint32_t* tmem_slot = (uint32_t*)(o_empty + O_STAGES);
    static_assert(O_STAGES % G == 0, "staging tiles rotate over the epilogue warp groups");
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t sbase = smem_u32(smem);
    const uint32_t rank = MC ? cluster_ctarank() : 0u;
    constexpr int TM = MC ? 2 * BM : BM;      // rows of a work tile (MC: one 128-row half per CTA)

    if (warp == 0 && lane == 0) {
        mbar_init(x_full, 1); mbar_init(x_empty, 1);
        for (int s = 0; s < F_UNITS; ++s) { mbar_init(&b_full[s], 1); mbar_init(&b_empty[s], MC ? 2 : 1); }
        for (int s = 0; s < HAND; ++s) mbar_init(&a_full[s], EPI_W / G);
        for (int s = 0; s < 3; ++s) mbar_init(&accf[s], 1);
        mbar_init(drained, EPI_W);
        for (int s = 0; s < O_STAGES; ++s) { mbar_init(&o_full[s], EPI_W / G); mbar_init(&o_empty[s], 1); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(512u));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    asm volatile("tcgen05.fence::before_thread_sync;");
    if (MC) cluster_sync_all(); else __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;");
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 2) {
        // ===== TMA stores: every staged [128 x 32] output tile -> h1 / h2 (words) / h3 (fp32) =====
        if (lane == 0) {
            uint32_t oi = 0;
            TileSeq seq(P.net[0].rows, P.net[1].rows, MC);
            int ni, tile;
            while (seq.next(ni, tile)) {
                const FwdNet& N = P.net[ni];
                for (int layer = 0; layer < 3; ++layer) {
                    const CUtensorMap* map = layer == 0 ? &N.mH1 : (layer == 1 ? &N.mH2 : &N.mH3);
                    const int nkb = (layer == 0 ? 256 : (layer == 1 ? N.n2 : 128)) / BK;
                    for (int kb = 0; kb < nkb; ++kb, ++oi) {
                        const uint32_t s = oi % O_STAGES, ph = (oi / O_STAGES) & 1;
                        mbar_wait_wd(&o_full[s], ph, 400 + (int)s);
                        tma_store_2d(map, sbase + F_OUT + s * KB_BYTES, kb * BK, tile * TM + (int)rank * BM);
                        // all but the newest store have finished READING their staging tile: hand the previous one back
                        asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(O_STAGES - 1) : "memory");
                        if (oi >= O_STAGES - 1) mbar_arrive(&o_empty[(oi - (O_STAGES - 1)) % O_STAGES]);
                    }
                }
            }
            asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
        }
    } else if (warp == 0) {
        // ===== TMA producer: the tile's input words, then the weight k-blocks of the three layers in consumption order =====
        if (lane == 0) {
            uint32_t u = 0, t = 0;
            auto wide = [&](const CUtensorMap* ma, const CUtensorMap* mb, int kb) {   // [256 x 32] B' and B'': one unit each
                for (int half = 0; half < 2; ++half, ++u) {
                    const uint32_t s = u % F_UNITS, ph = (u / F_UNITS) & 1;
                    mbar_wait_wd(&b_empty[s], ph ^ 1, 100 + (int)s);
                    if (H2_SKIP_BPP && half) { mbar_expect_tx(&b_full[s], 0); continue; }
                    mbar_expect_tx(&b_full[s], UNIT_BYTES);
                    load_rows<MC>(half ? mb : ma, &b_full[s], sbase + F_B + s * UNIT_BYTES, kb, 256, rank);
                }
            };
            auto narrow = [&](const CUtensorMap* ma, const CUtensorMap* mb, int kb) {   // B' + B'' of [128 x 32] in one unit
                const uint32_t s = u % F_UNITS, ph = (u / F_UNITS) & 1;
                mbar_wait_wd(&b_empty[s], ph ^ 1, 110 + (int)s);
                mbar_expect_tx(&b_full[s], H2_SKIP_BPP ? UNIT_BYTES / 2 : UNIT_BYTES);
                load_rows<MC>(ma, &b_full[s], sbase + F_B + s * UNIT_BYTES, kb, 128, rank);
                if (!H2_SKIP_BPP) load_rows<MC>(mb, &b_full[s], sbase + F_B + s * UNIT_BYTES + UNIT_BYTES / 2, kb, 128, rank);
                ++u;
            };
            TileSeq seq(P.net[0].rows, P.net[1].rows, MC);
            int ni, tile;
            while (seq.next(ni, tile)) {
                const FwdNet& N = P.net[ni];
                const int m0 = tile * TM + (int)rank * BM;
                mbar_wait_wd(x_empty, (t & 1) ^ 1, 120);
                mbar_expect_tx(x_full, 2 * KB_BYTES);
                tma_load_2d(&N.mX, x_full, sbase + F_X, 0, m0);
                tma_load_2d(&N.mX, x_full, sbase + F_X + KB_BYTES, BK, m0);
                for (int kb = 0; kb < 2; ++kb) wide(&N.mW1a, &N.mW1b, kb);
                for (int kb = 0; kb < 8; ++kb) { if (N.n2 == 256) wide(&N.mW2a, &N.mW2b, kb); else narrow(&N.mW2a, &N.mW2b, kb); }
                for (int kb = 0; kb < N.n2 / BK; ++kb) narrow(&N.mW3a, &N.mW3b, kb);
                ++t;
            }
        }
    } else if (warp == 1) {
        // ===== MMA issuer =====
        if (lane == 0) {
            uint32_t u = 0, t = 0, li = 0;
            CTL_DECL;
            // weight k-block of an N-row layer: waits for its unit(s), returns the operand addresses and the barriers to release
            auto weights = [&](int nn, uint32_t& b_a, uint32_t& b_b, uint64_t*& e0, uint64_t*& e1, int code) {
                e1 = nullptr;
                if (nn == 256) {
                    const uint32_t sa = u % F_UNITS, pa = (u / F_UNITS) & 1, sb = (u + 1) % F_UNITS, pb = ((u + 1) / F_UNITS) & 1;
                    u += 2;
                    CTL_WAIT(2, mbar_wait_wd(&b_full[sa], pa, code));
                    CTL_WAIT(2, mbar_wait_wd(&b_full[sb], pb, code + 1));
                    b_a = sbase + F_B + sa * UNIT_BYTES; b_b = sbase + F_B + sb * UNIT_BYTES;
                    e0 = &b_empty[sa]; e1 = &b_empty[sb];
                } else {
                    const uint32_t sa = u % F_UNITS, pa = (u / F_UNITS) & 1;
                    u += 1;
                    CTL_WAIT(2, mbar_wait_wd(&b_full[sa], pa, code + 2));
                    b_a = sbase + F_B + sa * UNIT_BYTES; b_b = b_a + UNIT_BYTES / 2;
                    e0 = &b_empty[sa];
                }
            };
            TileSeq seq(P.net[0].rows, P.net[1].rows, MC);
            int ni, tile;
            while (seq.next(ni, tile)) {
                const int n2 = P.net[ni].n2;
                const uint32_t par = t & 1;
                const uint32_t c1 = tmem_base + par * 256, c2 = tmem_base + (1 - par) * 256, c3 = c1;
                // ---- layer 1 (A = X words from shared memory)
                CTL_WAIT(0, mbar_wait_wd(x_full, par, 200));
                if (t > 0) CTL_WAIT(1, mbar_wait_wd(&accf[2], (t - 1) & 1, 201));   // layer 3 of the previous tile has finished READING h2 from c1's half
                asm volatile("tcgen05.fence::after_thread_sync;");
                {
                    // K = 64 is two k-blocks, both resident: the 8 small cross-term MMAs of BOTH k-blocks first and the 8 dominant ones last
                    // (TMEM accumulation truncates by up to an ulp of the accumulator on every add: the small terms land while it is small)
                    uint32_t b_a[2], b_b[2];
                    uint64_t *e0[2], *e1[2];
                    for (int kb = 0; kb < 2; ++kb) weights(256, b_a[kb], b_b[kb], e0[kb], e1[kb], 210);
                    asm volatile("tcgen05.fence::after_thread_sync;");
                    constexpr uint32_t id = idesc_f16_k(256);
                    // (descriptors: one per operand tile, + 2 per 32-byte k-step in the start-address field - the issuing thread's own
                    // instructions between two tcgen05.mma are on the critical path of the tensor pipe)
                    uint64_t da[2], dba[2], dbb[2];
#pragma unroll
                    for (int kb = 0; kb < 2; ++kb) { da[kb] = desc_kmajor(sbase + F_X + kb * KB_BYTES); dba[kb] = desc_kmajor(b_a[kb]); dbb[kb] = desc_kmajor(b_b[kb]); }
#pragma unroll
                    for (int kb = 0; kb < 2; ++kb) {
#pragma unroll
                        for (int k = 0; k < 4; ++k) umma_f16(c1, da[kb] + 2 * k, dbb[kb] + 2 * k, id, (kb | k) ? 1u : 0u);
                    }
#pragma unroll
                    for (int kb = 0; kb < 2; ++kb) {
#pragma unroll
                        for (int k = 0; k < 4; ++k) umma_f16(c1, da[kb] + 2 * k, dba[kb] + 2 * k, id, 1u);
                    }
                    for (int kb = 0; kb < 2; ++kb) {
                        release_unit<MC>(e0[kb]);
                        if (e1[kb]) release_unit<MC>(e1[kb]);
                    }
                }
                umma_commit(x_empty);
                umma_commit(&accf[0]);
                // ---- layers 2 and 3 (A words from TMEM in place)
                for (int layer = 0; layer < 2; ++layer) {
                    const int nk = layer == 0 ? 8 : n2 / BK, nn = layer == 0 ? n2 : 128;
                    const uint32_t ca = layer == 0 ? c1 : c2, cd = layer == 0 ? c2 : c3;
                    constexpr uint32_t id = idesc_f16_k(256);
                    // Measured (tools/micro/mma_rate.cu): a tcgen05.mma costs ~216 clk whatever its N, so every MMA here is N = 256.  A 128-wide
                    // layer's unit holds B' (rows 0-127) and B'' (rows 128-255) back to back = ONE [256 x 32] operand: a single MMA leaves the
                    // main products in columns [cd, cd + 128) and the cross terms in [cd + 128, cd + 256), which the epilogue adds (round to
                    // nearest) - half the MMAs, and the dominant sum is never rounded together with the small terms.
                    // layer 2 writes where the previous tile's layer-3 accumulator was: all 16 epilogue warps must have read it (the first
                    // hand-over below only proves that for ONE warp group)
                    if (layer == 0 && t > 0) CTL_WAIT(1, mbar_wait_wd(drained, (t - 1) & 1, 225));
                    for (int kb = 0; kb < nk; ++kb, ++li) {
                        uint32_t b_a, b_b;
                        uint64_t *e0, *e1;
                        weights(nn, b_a, b_b, e0, e1, 220);
                        CTL_WAIT(3 + layer, mbar_wait_wd(&a_full[li % HAND], (li / HAND) & 1, 230 + layer));   // the epilogue has written this k-block's words to TMEM
                        asm volatile("tcgen05.fence::after_thread_sync;");
                        const uint64_t dba = desc_kmajor(b_a), dbb = desc_kmajor(b_b);
                        const uint32_t a_t = ca + kb * BK;
                        if (nn == 128) {
#pragma unroll
                            for (int k = 0; k < 4; ++k) umma_f16_ts(cd, a_t + k * 8, dba + 2 * k, id, (kb | k) ? 1u : 0u);
                        } else {
#pragma unroll
                            for (int k = 0; k < 4; ++k) {
                                umma_f16_ts(cd, a_t + k * 8, dbb + 2 * k, id, (kb | k) ? 1u : 0u);
                                umma_f16_ts(cd, a_t + k * 8, dba + 2 * k, id, 1u);
                            }
                        }
                        release_unit<MC>(e0);
                        if (e1) release_unit<MC>(e1);
                    }
                    umma_commit(&accf[1 + layer]);
                }
                ++t;
            }
            CTL_FLUSH(0, 0, 5);
        }
    } else {
        // ===== epilogue warps: TMEM lane quarter q (rows); group grp takes the k-blocks kb = grp (mod G), slice sl of a k-block's columns =====
        const int q = warp & 3, eg = (warp - F_EPI0) >> 2, grp = eg % G, sl = eg / G;
        const uint32_t lane_off = (uint32_t)(q * 32) << 16;
        const int rit = q * 32 + lane;                 // row in this CTA's 128-row tile
        uint32_t t = 0, li_base = 0, oi_base = 0;
        CTL_DECL;
        TileSeq seq(P.net[0].rows, P.net[1].rows, MC);
        int ni, tile;
        while (seq.next(ni, tile)) {
            const FwdNet& N = P.net[ni];
            const int n2 = N.n2;
            const uint32_t par = t & 1;
            const uint32_t c1 = tmem_base + par * 256 + lane_off, c2 = tmem_base + (1 - par) * 256 + lane_off, c3 = c1;
#pragma unroll 1
            for (int layer = 0; layer < 3; ++layer) {
                const int width = layer == 0 ? 256 : (layer == 1 ? n2 : 128);
                const uint32_t cacc = layer == 0 ? c1 : (layer == 1 ? c2 : c3);
                const float* bias = layer == 0 ? N.b1s : (layer == 1 ? N.b2s : N.b3);
                const float inv = (layer == 0 ? 1.0f / (h2::S_X * h2::S_W) : 1.0f / (h2::S_ACT * h2::S_W)) * N.gain[layer];
                CTL_WAIT(layer, mbar_wait_wd(&accf[layer], par, 300 + layer));
                asm volatile("tcgen05.fence::after_thread_sync;");
                // software pipeline over the 8-column items, two register sets (unrolled by two, no copies): the accumulator slice and the
                // bias of item i + 1 are requested before item i is processed (tcgen05.ld is asynchronous until tcgen05.wait::ld)
                const bool two_acc = layer > 0 && width == 128;   // main | cross halves of the stacked MMA
                const int nkb = width / BK, nit = nkb;   // (nkb / G) k-blocks x G chunks: 8 or 4, always even
                auto colof = [&](int i) { return (grp + (i / G) * G) * BK + (sl * G + (i % G)) * 8; };
                auto request = [&](int i, uint32_t* r, uint32_t* r2, float4& b0, float4& b1) {
                    const int c = colof(i);
                    tmem_ld8_nowait(cacc + c, r);
                    if (two_acc) tmem_ld8_nowait(cacc + 128 + c, r2);
                    b0 = __ldg(reinterpret_cast<const float4*>(bias + c)); b1 = __ldg(reinterpret_cast<const float4*>(bias + c + 4));
                };
                auto process = [&](int i, const uint32_t* r, const uint32_t* r2, const float4& b0, const float4& b1) {
                    const int kb = grp + (i / G) * G, c8 = sl * G + (i % G), col = kb * BK + c8 * 8;
                    const float bb[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
                    const uint32_t oi = oi_base + (uint32_t)kb, os = oi % O_STAGES;
                    if ((i % G) == 0) CTL_WAIT(4, mbar_wait_wd(&o_empty[os], ((oi / O_STAGES) & 1) ^ 1, 320));   // the staging tile is free
                    uint32_t w[8];
                    if (layer < 2) {
                        // h2 word of 16 ELU(x), x = acc inv + b, computed on y = 16 x (bias pre-scaled): 16 (e^x - 1) = 16 ex2(y log2(e) / 16) - 16
#pragma unroll
                        for (int j = 0; j < 8; ++j) {
                            const float a = two_acc ? __uint_as_float(r[j]) + __uint_as_float(r2[j]) : __uint_as_float(r[j]);
                            const float y = fmaf(a, inv * h2::S_ACT, bb[j]);
                            float e;
                            asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(y * (1.4426950408889634f / h2::S_ACT)));
                            w[j] = h2::pack(fminf((y > 0.0f) ? y : fmaf(e, h2::S_ACT, -h2::S_ACT), h2::H2_MAX));
                        }
                        tmem_st8(cacc + col, w);
                        if ((i % G) == G - 1) {   // the k-block's last chunk: hand it over
                            const uint32_t li = li_base + (uint32_t)kb;
                            CTL_WAIT(6, handover(&a_full[li % HAND], lane));
                        }
                    } else {
#pragma unroll
                        for (int j = 0; j < 8; ++j) w[j] = __float_as_uint(elu_fast5(fmaf(__uint_as_float(r[j]) + __uint_as_float(r2[j]), inv, bb[j])));
                    }
                    // the HBM copy (after the hand-over): staging tile -> bulk tensor store by warp 2
                    stage_words(sbase + F_OUT + os * KB_BYTES, rit, c8, w);
                    if ((i % G) == G - 1) stage_done(&o_full[os], lane);
                };
                uint32_t ra[8], ra2[8], rb[8], rb2[8];
                float4 ba0, ba1, bb0, bb1;
                request(0, ra, ra2, ba0, ba1);
#pragma unroll 1
                for (int i = 0; i < nit; i += 2) {
                    if (two_acc) tmem_wait_ld16(ra, ra2); else tmem_wait_ld8(ra);
                    request(i + 1, rb, rb2, bb0, bb1);
                    process(i, ra, ra2, ba0, ba1);
                    if (two_acc) tmem_wait_ld16(rb, rb2); else tmem_wait_ld8(rb);
                    if (i + 2 < nit) request(i + 2, ra, ra2, ba0, ba1);
                    process(i + 1, rb, rb2, bb0, bb1);
                }
                if (layer < 2) li_base += (uint32_t)nkb;
                oi_base += (uint32_t)nkb;
            }
            asm volatile("tcgen05.fence::before_thread_sync;");
            __syncwarp();
            if (lane == 0) mbar_arrive(drained);
            ++t;
        }
        if (threadIdx.x == 32 * F_EPI0) CTL_FLUSH(0, 8, 7);
    }
    asm volatile("tcgen05.fence::before_thread_sync;");
    if (MC) cluster_sync_all(); else __syncthreads();   // (MC: the peer's multicasts / commits into this CTA have all landed)
    asm volatile("tcgen05.fence::after_thread_sync;");
    if (warp == 1) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u));
}

// =====================================================================================================================
// backward (input gradients of the hidden layers + bias gradients)
// =====================================================================================================================
struct alignas(64) BwdNet {
    CUtensorMap mZ3, mH2, mH1;            // words: dz3 [rows,128], h2 [rows,n2], h1 [rows,256]; box 32 x 128 (the epilogue's "aux" tiles)
    CUtensorMap mW3Ta, mW3Tb;             // W3^T [n2, 128]   box 32 x n2
    CUtensorMap mW2Ta, mW2Tb;             // W2^T [256, n2]   box 32 x 256
    CUtensorMap mDZ2, mDZ1;               // outputs: words [rows, n2], [rows, 256] (scaled by the net's gradient scale); box 32 x 128
    float *db2, *db1;                     // bias gradients of layers 2 and 1 (+= column sums of dz2 / dz1)
    const float* isg;                     // device: 1 / gradient scale of this net
    int rows, n2;
    float gain[2];                        // midpoint compensation of the dz2 / dz1 accumulators (see FwdNet)
};
struct alignas(64) BwdParams { BwdNet net[2]; };

#ifndef H2_B_UNITS
#define H2_B_UNITS 4
#define H2_B_AUX 4
#endif
static constexpr int B_UNITS = H2_B_UNITS, B_AUX_STAGES = H2_B_AUX;
static constexpr int B_EPI0 = 4;                           // warp 0 weight TMA, 1 MMA, 2 aux TMA, 3 TMA stores
static constexpr int B_THREADS = 32 * (B_EPI0 + EPI_W);
static constexpr int B_B = 0, B_AUX = B_UNITS * UNIT_BYTES, B_OUT = B_AUX + B_AUX_STAGES * KB_BYTES, B_BAR = B_OUT + O_STAGES * KB_BYTES,
                     B_SMEM = B_BAR + 256 + 1024;
static_assert(B_SMEM <= 232448, "shared memory budget");

// column sums of this warp's 32 rows x 8 columns -> atomics on dst[0..7], scaled
__device__ __forceinline__ void colsum8s(float* v, int lane, float* dst, float scale) {
#pragma unroll
    for (int half = 4; half >= 1; half >>= 1) {
        const bool upper = (lane & half) != 0;
#pragma unroll
        for (int j = 0; j < half; ++j) {
            const float send = upper ? v[j] : v[j + half];
            const float keep = upper ? v[j + half] : v[j];
            v[j] = keep + __shfl_xor_sync(0xffffffffu, send, half);
        }
    }
    float tot = v[0] + __shfl_xor_sync(0xffffffffu, v[0], 8);
    tot += __shfl_xor_sync(0xffffffffu, tot, 16);
    if (lane < 8) atomicAdd(dst + lane, tot * scale);   // lane l holds column l (bit i of l picked the upper half at step 2^i)
}

template <int G, int MC>   // epilogue warp groups, weight multicast in CTA pairs (see k_mlp_fwd_h2)
__global__ void __launch_bounds__(B_THREADS, 1) k_mlp_bwd_h2(const __grid_constant__ BwdParams P) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    uint64_t* b_full = (uint64_t*)(smem + B_BAR);
    uint64_t* b_empty = b_full + B_UNITS;
    uint64_t* a_full = b_empty + B_UNITS;           // [HAND]
    uint64_t* aux_full = a_full + HAND;
    uint64_t* aux_empty = aux_full + B_AUX_STAGES;
    uint64_t* accf = aux_empty + B_AUX_STAGES;   // [2]: dh2 complete, dh1 complete
    uint64_t* drained = accf + 2;                // every epilogue warp has read the tile's dh1 (the next tile's dh1 goes to the same columns)
    uint64_t* o_full = drained + 1;              // [O_STAGES] output staging tiles (see k_mlp_fwd_h2)
    uint64_t* o_empty = o_full + O_STAGES;
    uint32_t* tmem_slot = (uint32_t*)(o_empty + O_STAGES);
    static_assert(O_STAGES % G == 0, "staging tiles rotate over the epilogue warp groups");
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t sbase = smem_u32(smem);
    const uint32_t rank = MC ? cluster_ctarank() : 0u;
    constexpr int TM = MC ? 2 * BM : BM;

    if (warp == 0 && lane == 0) {
        for (int s = 0; s < B_UNITS; ++s) { mbar_init(&b_full[s], 1); mbar_init(&b_empty[s], MC ? 2 : 1); }
        for (int s = 0; s < HAND; ++s) mbar_init(&a_full[s], EPI_W / G);
        for (int s = 0; s < B_AUX_STAGES; ++s) { mbar_init(&aux_full[s], 1); mbar_init(&aux_empty[s], EPI_W / G); }
        mbar_init(&accf[0], 1); mbar_init(&accf[1], 1);
        mbar_init(drained, EPI_W);
        for (int s = 0; s < O_STAGES; ++s) { mbar_init(&o_full[s], EPI_W / G); mbar_init(&o_empty[s], 1); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(512u));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    asm volatile("tcgen05.fence::before_thread_sync;");
    if (MC) cluster_sync_all(); else __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;");
    const uint32_t tmem_base = *tmem_slot;
    // TMEM columns: dz3 words [0,128); dh2 / dz2 [256, 512) (n2 = 128: main products [256,384) | cross terms [384,512)); dh1 [0,256) - its first
    // half aliases dz3 (dead once dh2 is complete): the epilogue drains that half first, stages the NEXT tile's dz3 there, then drains the rest
    constexpr uint32_t CZ = 0, CA = 256, CD = 0;

    if (warp == 3) {
        // ===== TMA stores: the staged [128 x 32] tiles of dz2, then dz1, per row tile =====
        if (lane == 0) {
            uint32_t oi = 0;
            TileSeq seq(P.net[0].rows, P.net[1].rows, MC);
            int ni, tile;
            while (seq.next(ni, tile)) {
                const BwdNet& N = P.net[ni];
                for (int seg = 0; seg < 2; ++seg) {
                    const CUtensorMap* map = seg == 0 ? &N.mDZ2 : &N.mDZ1;
                    const int nkb = (seg == 0 ? N.n2 : 256) / BK;
                    for (int kb = 0; kb < nkb; ++kb, ++oi) {
                        const uint32_t s = oi % O_STAGES, ph = (oi / O_STAGES) & 1;
                        mbar_wait_wd(&o_full[s], ph, 800 + (int)s);
                        tma_store_2d(map, sbase + B_OUT + s * KB_BYTES, kb * BK, tile * TM + (int)rank * BM);
                        asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(O_STAGES - 1) : "memory");
                        if (oi >= O_STAGES - 1) mbar_arrive(&o_empty[(oi - (O_STAGES - 1)) % O_STAGES]);
                    }
                }
            }
            asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
        }
    } else if (warp == 0) {
        // ===== TMA producer: weight k-blocks (W3^T then W2^T per tile) =====
        if (lane == 0) {
            uint32_t u = 0;
            auto wide = [&](const CUtensorMap* ma, const CUtensorMap* mb, int kb) {
                for (int half = 0; half < 2; ++half, ++u) {
                    const uint32_t s = u % B_UNITS, ph = (u / B_UNITS) & 1;
                    mbar_wait_wd(&b_empty[s], ph ^ 1, 500 + (int)s);
                    mbar_expect_tx(&b_full[s], UNIT_BYTES);
                    load_rows<MC>(half ? mb : ma, &b_full[s], sbase + B_B + s * UNIT_BYTES, kb, 256, rank);
                }
            };
            auto narrow = [&](const CUtensorMap* ma, const CUtensorMap* mb, int kb) {
                const uint32_t s = u % B_UNITS, ph = (u / B_UNITS) & 1;
                mbar_wait_wd(&b_empty[s], ph ^ 1, 510 + (int)s);
                mbar_expect_tx(&b_full[s], UNIT_BYTES);
                load_rows<MC>(ma, &b_full[s], sbase + B_B + s * UNIT_BYTES, kb, 128, rank);
                load_rows<MC>(mb, &b_full[s], sbase + B_B + s * UNIT_BYTES + UNIT_BYTES / 2, kb, 128, rank);
                ++u;
            };
            TileSeq seq(P.net[0].rows, P.net[1].rows, MC);
            int ni, tile;
            while (seq.next(ni, tile)) {
                const BwdNet& N = P.net[ni];
                for (int kb = 0; kb < 4; ++kb) { if (N.n2 == 256) wide(&N.mW3Ta, &N.mW3Tb, kb); else narrow(&N.mW3Ta, &N.mW3Tb, kb); }
                for (int kb = 0; kb < N.n2 / BK; ++kb) wide(&N.mW2Ta, &N.mW2Tb, kb);
            }
        }
    } else if (warp == 2) {
        // ===== TMA producer: aux tiles (this CTA's 128 rows) in the epilogue's consumption order =====
        if (lane == 0) {
            uint32_t ai = 0;
            auto aux = [&](const CUtensorMap* m, int kb, int m0) {
                const uint32_t s = ai % B_AUX_STAGES, ph = (ai / B_AUX_STAGES) & 1;
                mbar_wait_wd(&aux_empty[s], ph ^ 1, 520 + (int)s);
                mbar_expect_tx(&aux_full[s], KB_BYTES);
                tma_load_2d(m, &aux_full[s], sbase + B_AUX + s * KB_BYTES, kb * BK, m0);
                ++ai;
            };
            TileSeq seq(P.net[0].rows, P.net[1].rows, MC);
            int cn, ct, nn = 0, nt = 0;
            bool have = seq.next(cn, ct);
            if (have) for (int kb = 0; kb < 4; ++kb) aux(&P.net[cn].mZ3, kb, ct * TM + (int)rank * BM);
            while (have) {
                const bool hn = seq.next(nn, nt);
                const BwdNet& N = P.net[cn];
                const int m0 = ct * TM + (int)rank * BM;
                for (int kb = 0; kb < N.n2 / BK; ++kb) aux(&N.mH2, kb, m0);
                for (int kb = 0; kb < 4; ++kb) aux(&N.mH1, kb, m0);
                if (hn) for (int kb = 0; kb < 4; ++kb) aux(&P.net[nn].mZ3, kb, nt * TM + (int)rank * BM);
                for (int kb = 4; kb < 8; ++kb) aux(&N.mH1, kb, m0);
                have = hn; cn = nn; ct = nt;
            }
        }
    } else if (warp == 1) {
        // ===== MMA issuer =====
        if (lane == 0) {
            uint32_t u = 0, li = 0, t = 0;
            CTL_DECL;
            auto weights = [&](bool two_units, uint32_t& b_a, uint32_t& b_b, uint64_t*& e0, uint64_t*& e1, int code) {
                e1 = nullptr;
                if (two_units) {
                    const uint32_t sa = u % B_UNITS, pa = (u / B_UNITS) & 1, sb = (u + 1) % B_UNITS, pb = ((u + 1) / B_UNITS) & 1;
                    u += 2;
                    CTL_WAIT(2, mbar_wait_wd(&b_full[sa], pa, code));
                    CTL_WAIT(2, mbar_wait_wd(&b_full[sb], pb, code + 1));
                    b_a = sbase + B_B + sa * UNIT_BYTES; b_b = sbase + B_B + sb * UNIT_BYTES;
                    e0 = &b_empty[sa]; e1 = &b_empty[sb];
                } else {
                    const uint32_t sa = u % B_UNITS, pa = (u / B_UNITS) & 1;
                    u += 1;
                    CTL_WAIT(2, mbar_wait_wd(&b_full[sa], pa, code + 2));
                    b_a = sbase + B_B + sa * UNIT_BYTES; b_b = b_a + UNIT_BYTES / 2;
                    e0 = &b_empty[sa];
                }
            };
            TileSeq seq(P.net[0].rows, P.net[1].rows, MC);
            int ni, tile;
            while (seq.next(ni, tile)) {
                const int n2 = P.net[ni].n2;
                // ---- dh2 [128, n2] = dz3 [128,128] W3: 4 k-blocks
                for (int kb = 0; kb < 4; ++kb, ++li) {
                    uint32_t b_a, b_b;
                    uint64_t *e0, *e1;
                    weights(n2 == 256, b_a, b_b, e0, e1, 600);
                    CTL_WAIT(3, mbar_wait_wd(&a_full[li % HAND], (li / HAND) & 1, 610));
                    asm volatile("tcgen05.fence::after_thread_sync;");
                    constexpr uint32_t id = idesc_f16_k(256);   // (every MMA is N = 256: see k_mlp_fwd_h2)
                    const uint64_t dba = desc_kmajor(b_a), dbb = desc_kmajor(b_b);
                    const uint32_t a_t = tmem_base + CZ + kb * BK, cd = tmem_base + CA;
                    if (n2 == 128) {   // stacked [B'; B''] unit: main | cross
#pragma unroll
                        for (int k = 0; k < 4; ++k) umma_f16_ts(cd, a_t + k * 8, dba + 2 * k, id, (kb | k) ? 1u : 0u);
                    } else {
#pragma unroll
                        for (int k = 0; k < 4; ++k) {
                            umma_f16_ts(cd, a_t + k * 8, dbb + 2 * k, id, (kb | k) ? 1u : 0u);
                            umma_f16_ts(cd, a_t + k * 8, dba + 2 * k, id, 1u);
                        }
                    }
                    release_unit<MC>(e0);
                    if (e1) release_unit<MC>(e1);
                }
                umma_commit(&accf[0]);
                // ---- dh1 [128, 256] = dz2 [128, n2] W2: n2 / 32 k-blocks
                if (t > 0) CTL_WAIT(1, mbar_wait_wd(drained, (t - 1) & 1, 625));   // all 16 epilogue warps have read the previous tile's dh1
                for (int kb = 0; kb < n2 / BK; ++kb, ++li) {
                    uint32_t b_a, b_b;
                    uint64_t *e0, *e1;
                    weights(true, b_a, b_b, e0, e1, 620);
                    CTL_WAIT(4, mbar_wait_wd(&a_full[li % HAND], (li / HAND) & 1, 630));
                    asm volatile("tcgen05.fence::after_thread_sync;");
                    constexpr uint32_t id = idesc_f16_k(256);
                    const uint64_t dba = desc_kmajor(b_a), dbb = desc_kmajor(b_b);
                    const uint32_t a_t = tmem_base + CA + kb * BK, cd = tmem_base + CD;
#pragma unroll
                    for (int k = 0; k < 4; ++k) {
                        umma_f16_ts(cd, a_t + k * 8, dbb + 2 * k, id, (kb | k) ? 1u : 0u);
                        umma_f16_ts(cd, a_t + k * 8, dba + 2 * k, id, 1u);
                    }
                    release_unit<MC>(e0);
                    if (e1) release_unit<MC>(e1);
                }
                umma_commit(&accf[1]);
                ++t;
            }
            CTL_FLUSH(1, 0, 5);
        }
    } else if (warp >= B_EPI0) {
        // ===== epilogue warps: group grp takes the k-blocks kb = grp (mod G) of every segment, G chunks of 8 columns per thread =====
        const int q = warp & 3, eg = (warp - B_EPI0) >> 2, grp = eg % G, sl = eg / G;
        const uint32_t lane_off = (uint32_t)(q * 32) << 16;
        const int rit = q * 32 + lane;
        const uint32_t sw = (uint32_t)(rit & 7);
        uint32_t t = 0, li_base = 0, ai_base = 0, oi_base = 0;   // hand-overs / aux tiles / output tiles before the current segment (every segment is a multiple of G long)
        CTL_DECL;
        // this thread's 8 words (columns c8 * 8 ...) of the aux k-block in ring stage s (128-byte swizzled rows)
        auto aux_wait = [&](uint32_t ai) -> uint32_t {
            const uint32_t s = ai % B_AUX_STAGES, ph = (ai / B_AUX_STAGES) & 1;
            CTL_WAIT(2, mbar_wait_wd(&aux_full[s], ph, 700));
            return s;
        };
        auto aux_read = [&](uint32_t s, int c8, uint32_t* h) {
            const uint32_t rbase = sbase + B_AUX + s * KB_BYTES + (uint32_t)rit * 128u;
            const float4 x0 = lds_v4(rbase + ((((uint32_t)(2 * c8)) ^ sw) << 4)), x1 = lds_v4(rbase + ((((uint32_t)(2 * c8 + 1)) ^ sw) << 4));
            h[0] = __float_as_uint(x0.x); h[1] = __float_as_uint(x0.y); h[2] = __float_as_uint(x0.z); h[3] = __float_as_uint(x0.w);
            h[4] = __float_as_uint(x1.x); h[5] = __float_as_uint(x1.y); h[6] = __float_as_uint(x1.z); h[7] = __float_as_uint(x1.w);
        };
        // The stage is handed back to the TMA producer only AFTER instructions that consume the loaded registers have issued: an mbarrier
        // arrive right behind the ld.shared can overtake the loads still in flight (see mlp_chain.cuh).
        auto aux_release = [&](uint32_t s) {
            __syncwarp();
            if (lane == 0) mbar_arrive(&aux_empty[s]);
        };
        auto stage_dz3 = [&]() {   // dz3 of a tile: aux tiles -> A operand, the words as they are
#pragma unroll 1
            for (int kb = grp; kb < 4; kb += G) {
                const uint32_t as = aux_wait(ai_base + kb);
#pragma unroll
                for (int c = 0; c < G; ++c) {
                    const int c8 = sl * G + c;
                    uint32_t w[8];
                    aux_read(as, c8, w);
                    tmem_st8(tmem_base + lane_off + CZ + kb * BK + c8 * 8, w);
                }
                handover(&a_full[(li_base + kb) % HAND], lane);
                aux_release(as);
            }
            ai_base += 4; li_base += 4;
        };
        // a segment of ELU'-masked gradient k-blocks: v = acc * ELU'(h); store; column sums; optionally hand over to the next GEMM
        auto grad_segment = [&](int nkb, uint32_t tcol0, uint32_t cross, int col0, float* db, float isg, bool row_ok, bool handoff, float gain) {
            const float ascale = (1.0f / h2::S_W) * gain;
#pragma unroll 1
            for (int kb = grp; kb < nkb; kb += G) {
                const uint32_t as = aux_wait(ai_base + kb);
                const uint32_t oi = oi_base + (uint32_t)kb, os = oi % O_STAGES;
                CTL_WAIT(3, mbar_wait_wd(&o_empty[os], ((oi / O_STAGES) & 1) ^ 1, 730));   // the staging tile is free
#pragma unroll 1
                for (int c = 0; c < G; ++c) {
                    const int c8 = sl * G + c;
                    const uint32_t tcol = tcol0 + kb * BK + c8 * 8;
                    const int col = col0 + kb * BK + c8 * 8;
                    uint32_t hw[8], r[8], r2[8];
                    aux_read(as, c8, hw);
                    tmem_ld8_nowait(tmem_base + lane_off + tcol, r);
                    if (cross) tmem_ld8_nowait(tmem_base + lane_off + tcol + cross, r2);
                    if (cross) tmem_wait_ld16(r, r2); else tmem_wait_ld8(r);
                    float v[8];
                    uint32_t w[8];
#pragma unroll
                    for (int j = 0; j < 8; ++j) {
                        const float h = h2::unpack(hw[j]);   // = S_ACT * activation
                        const float acc = cross ? __uint_as_float(r[j]) + __uint_as_float(r2[j]) : __uint_as_float(r[j]);
                        v[j] = acc * ascale * ((h > 0.0f) ? 1.0f : fmaf(h, 1.0f / h2::S_ACT, 1.0f));
                        w[j] = h2::pack(v[j]);
                    }
                    if (handoff) {
                        tmem_st8(tmem_base + lane_off + tcol, w);
                        if (c == G - 1) CTL_WAIT(3, handover(&a_full[(li_base + kb) % HAND], lane));
                    }
                    stage_words(sbase + B_OUT + os * KB_BYTES, rit, c8, w);
                    if (!row_ok) {
#pragma unroll
                        for (int j = 0; j < 8; ++j) v[j] = 0.0f;
                    }
                    colsum8s(v, lane, db + col, isg);
                }
                stage_done(&o_full[os], lane);
                aux_release(as);
            }
            ai_base += (uint32_t)nkb;
            oi_base += (uint32_t)nkb;
            if (handoff) li_base += (uint32_t)nkb;
        };
        TileSeq seq(P.net[0].rows, P.net[1].rows, MC);
        int cn, ct, nn = 0, nt = 0;
        bool have = seq.next(cn, ct);
        if (have) stage_dz3();
        while (have) {
            const bool hn = seq.next(nn, nt);
            const BwdNet& N = P.net[cn];
            const int n2 = N.n2;
            const bool row_ok = ct * TM + (int)rank * BM + rit < N.rows;
            const uint32_t par = t & 1;
            const float isg = __ldg(N.isg);
            // dz2
            CTL_WAIT(0, mbar_wait_wd(&accf[0], par, 720));
            asm volatile("tcgen05.fence::after_thread_sync;");
            grad_segment(n2 / BK, CA, n2 == 128 ? 128u : 0u, 0, N.db2, isg, row_ok, true, N.gain[0]);
            // dz1: the half that shares TMEM columns with the next tile's dz3 first
            CTL_WAIT(1, mbar_wait_wd(&accf[1], par, 721));
            asm volatile("tcgen05.fence::after_thread_sync;");
            grad_segment(4, CD, 0u, 0, N.db1, isg, row_ok, false, N.gain[1]);
            if (hn) stage_dz3();
            grad_segment(4, CD + 128, 0u, 128, N.db1, isg, row_ok, false, N.gain[1]);
            asm volatile("tcgen05.fence::before_thread_sync;");
            __syncwarp();
            if (lane == 0) mbar_arrive(drained);
            have = hn; cn = nn; ct = nt;
            ++t;
        }
        if (threadIdx.x == 32 * B_EPI0) CTL_FLUSH(1, 8, 4);
    }
    asm volatile("tcgen05.fence::before_thread_sync;");
    if (MC) cluster_sync_all(); else __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;");
    if (warp == 1) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u));
}

}  // namespace chain2
}  // namespace b200
