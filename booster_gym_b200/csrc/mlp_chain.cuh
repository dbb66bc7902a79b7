// mlp_chain.cuh - the three hidden layers of an actor / critic MLP (utils/model.py:9-26) as ONE persistent tcgen05 kernel per
// direction: a 128-row tile's activations never leave the SM between layers.
//
//   forward   X[128,64] -> h1 = ELU(X W1^T + b1) -> h2 = ELU(h1 W2^T + b2) -> h3 = ELU(h2 W3^T + b3)
//   backward  dz3[128,128] -> dz2 = (dz3 W3) * ELU'(h2) -> dz1 = (dz2 W2) * ELU'(h1)        (+ bias gradients = column sums)
//
// How a layer's output becomes the next layer's A operand without touching shared memory for the hi half:
//   the accumulator of layer l lives in TMEM (128 lanes = rows, one column per output unit).  The epilogue warps read a
//   32-column chunk (tcgen05.ld), apply bias + ELU (forward) or ELU'(h) (backward), store the fp32 result to HBM for the
//   other direction / the weight-gradient GEMMs, and write  hi = tf32_rna(v)  BACK IN PLACE (tcgen05.st): the accumulator's
//   columns ARE the K index of the next layer, so the next tcgen05.mma takes its A operand straight from TMEM ("TS" form,
//   tcgen05.mma [d], [a_tmem], b_desc).  Only the small half  lo = tf32(v - hi)  goes through a shared-memory ring (16 KB per
//   32-column k-block, 128-byte swizzle, written by the same epilogue threads).  Per k-step of 8:
//        D += A_lo(smem) * B_hi      D += A_hi(TMEM) * B_lo      D += A_hi(TMEM) * B_hi          (3xTF32, fp32-class products)
//   so layer l+1's MMAs trail layer l's epilogue chunk by chunk (mbarrier per chunk), and the tensor pipe stays busy while
//   the epilogue runs.  Weights (pre-split hi / lo by k_weight_prep) stream from L2 through a 4 x 32 KB TMA ring.
//
// TMEM map (512 columns, forward; p = tile parity of this CTA):  acc1/h1 -> [p*256, +256)   acc2/h2 -> [(1-p)*256, +n2)
//   acc3 -> [p*256, +128) (h1 is dead once layer 2 has completed).  Layer 1 of the next tile goes to the other half, so it
//   runs while the epilogue drains acc3.
// TMEM map (backward): dz3 -> [0,128)   dh2/dz2 -> [128, 128+n2)   dh1 -> low half of the columns in [384,512), high half
//   in [0,128) (dz3 is dead by then); the epilogue drains the high half first, then stages the NEXT tile's dz3 there, then
//   drains the low half, so the next tile's first GEMM overlaps the drain.
//
// HBM traffic per sample (critic, fp32): forward reads X hi/lo (512 B) and writes h1, h2, h3 (2 560 B) - the layer-by-layer
// path also READ h1 and h2 back (2 048 B) and paid a TMA -> shared-memory -> converter-warp pass per layer.  Backward reads
// dz3, h2, h1 (2 560 B) and writes dz2, dz1 (2 048 B); the layer-by-layer path re-read dz3, dz2 (1 536 B) on top.
#pragma once
#include "gemm_tc.cuh"

namespace b200 {
namespace chain {
using namespace tc;

static constexpr int UNIT_BYTES = 32768;          // one weight-ring unit: [256 x 32] fp32 (hi OR lo) or [128 x 32] hi + lo
static constexpr int NUNITS = 4;
static constexpr int KB_BYTES = BM * BK * 4;      // one [128 x 32] fp32 k-block tile

// Optional per-CTA wait accounting (-DB200_CHAIN_TL, tools/chain_timeline.py): SM-clock cycles each role spent blocked on each
// barrier class, per CTA.  Compiled out of the product build.
#ifdef B200_CHAIN_TL
__device__ long long g_chain_tl[2][160][16];
#define CTL_DECL long long ctl_[8] = {0, 0, 0, 0, 0, 0, 0, 0}; const long long ctl_t0_ = clock64()
#define CTL_WAIT(slot, stmt) do { const long long c0_ = clock64(); stmt; ctl_[slot] += clock64() - c0_; } while (0)
#define CTL_FLUSH(kernel, base, n) do { for (int i_ = 0; i_ < (n); ++i_) g_chain_tl[kernel][blockIdx.x < 160 ? blockIdx.x : 159][(base) + i_] = ctl_[i_]; \
                                        g_chain_tl[kernel][blockIdx.x < 160 ? blockIdx.x : 159][(base) + (n)] = clock64() - ctl_t0_; } while (0)
#else
#define CTL_DECL
#define CTL_WAIT(slot, stmt) stmt
#define CTL_FLUSH(kernel, base, n)
#endif

__device__ unsigned int g_chain_error = 0;        // last watchdog code (a barrier that never completed traps the kernel)

// bounded mbarrier wait: a protocol bug becomes a trapped kernel (CUDA error at the next sync), never a hung GPU
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
            atomicExch(&g_chain_error, (unsigned)code);
            printf("mlp_chain watchdog: block %d thread %d barrier code %d parity %u\n", (int)blockIdx.x, (int)threadIdx.x, code, parity);
            __trap();
        }
    } while (!done);
}

__device__ __forceinline__ void umma_tf32_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t db, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}"
        ::"r"(tmem_d), "r"(tmem_a), "l"(db), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void umma_tf32_ts_pair(uint32_t tmem_d, uint32_t tmem_a, uint64_t db, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::2.kind::tf32 [%0], [%1], %2, %3, p;\n\t}"
        ::"r"(tmem_d), "r"(tmem_a), "l"(db), "r"(idesc), "r"(accumulate) : "memory");
}
// PAIR = 1 (cta_group::2): a cluster of two CTAs (the two SMs of a TPC) works on 256 rows.  Each CTA keeps ITS 128 rows' accumulators
// / A operands in its own TMEM and lo ring and stages HALF of every weight k-block; the leader CTA issues the MMAs (M = 256), the
// tensor cores exchange the weight halves.  Per-CTA shared-memory traffic of the weight operand - the measured limiter of the
// single-CTA kernel (3 MMAs per k-step each re-read the whole tile: ~125-150 B/clk needed of 128 B/clk) - halves.
#ifdef B200_CHAIN_DRY   // measurement build: no MMAs are issued (commits complete at once) - what is left is the epilogue + TMA time
template <int PAIR> __device__ __forceinline__ void mma_ss(uint32_t, uint64_t, uint64_t, uint32_t, uint32_t) {}
template <int PAIR> __device__ __forceinline__ void mma_ts(uint32_t, uint32_t, uint64_t, uint32_t, uint32_t) {}
#else
template <int PAIR> __device__ __forceinline__ void mma_ss(uint32_t d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t acc) {
    if (PAIR) umma_tf32_pair(d, da, db, idesc, acc); else umma_tf32(d, da, db, idesc, acc);
}
template <int PAIR> __device__ __forceinline__ void mma_ts(uint32_t d, uint32_t a, uint64_t db, uint32_t idesc, uint32_t acc) {
    if (PAIR) umma_tf32_ts_pair(d, a, db, idesc, acc); else umma_tf32_ts(d, a, db, idesc, acc);
}
#endif
template <int PAIR> __device__ __forceinline__ void commit(uint64_t* bar) {   // PAIR: arrives on `bar` in BOTH CTAs
    if (PAIR) umma_commit_pair(bar); else umma_commit(bar);
}
// wait on a barrier of the leader CTA that also collects arrivals / TMA bytes of the peer CTA
template <int PAIR> __device__ __forceinline__ void wait_collect(uint64_t* b, uint32_t parity, int code) {
    if (!PAIR) { mbar_wait_wd(b, parity, code); return; }
    uint32_t done;
    const uint32_t a = smem_u32(b);
    const long long t0 = clock64();
    do {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(done) : "r"(a), "r"(parity) : "memory");
        if (!done && clock64() - t0 > 4000000000ll) {
            atomicExch(&g_chain_error, (unsigned)code);
            printf("mlp_chain watchdog: block %d thread %d cluster barrier code %d parity %u\n", (int)blockIdx.x, (int)threadIdx.x, code, parity);
            __trap();
        }
    } while (!done);
}
// producer side of such a barrier: the leader announces the bytes of BOTH CTAs, the peer only arrives
template <int PAIR> __device__ __forceinline__ void expect_collect(uint64_t* bar, uint32_t bytes_per_cta, uint32_t rank) {
    if (!PAIR) { mbar_expect_tx(bar, bytes_per_cta); return; }
    if (rank == 0) mbar_expect_tx(bar, 2 * bytes_per_cta); else mbar_arrive_leader(bar);
}
template <int PAIR> __device__ __forceinline__ void tma_collect(const CUtensorMap* map, uint64_t* bar, uint32_t dst, int c0, int c1) {
    if (PAIR) tma_load_2d_pair(map, bar, dst, c0, c1); else tma_load_2d(map, bar, dst, c0, c1);
}
__device__ __forceinline__ void tmem_ld8(uint32_t taddr, uint32_t* r) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]) : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_ld8_nowait(uint32_t taddr, uint32_t* r) {   // results are valid after tmem_wait_ld()
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]) : "r"(taddr));
}
__device__ __forceinline__ void tmem_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
// ... and the compiler must not read the registers of an outstanding tcgen05.ld before the wait: tie them to it
__device__ __forceinline__ void tmem_wait_ld8(uint32_t* r) {
    asm volatile("tcgen05.wait::ld.sync.aligned;" : "+r"(r[0]), "+r"(r[1]), "+r"(r[2]), "+r"(r[3]), "+r"(r[4]), "+r"(r[5]), "+r"(r[6]), "+r"(r[7]) :: "memory");
}
__device__ __forceinline__ void tmem_wait_ld16(uint32_t* r, uint32_t* q) {
    asm volatile("tcgen05.wait::ld.sync.aligned;" : "+r"(r[0]), "+r"(r[1]), "+r"(r[2]), "+r"(r[3]), "+r"(r[4]), "+r"(r[5]), "+r"(r[6]), "+r"(r[7]),
                 "+r"(q[0]), "+r"(q[1]), "+r"(q[2]), "+r"(q[3]), "+r"(q[4]), "+r"(q[5]), "+r"(q[6]), "+r"(q[7]) :: "memory");
}
__device__ __forceinline__ void tmem_st8(uint32_t taddr, const uint32_t* r) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};"
                 ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]) : "memory");
}
__device__ __forceinline__ void tmem_wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

// CTA b's share of the tiles of two nets: net 0 round-robin from the front, net 1 from the back, so the CTAs that got one
// more (expensive) net-0 tile get one less net-1 tile.  Every warp role walks the same sequence.
struct TileSeq {
    int nt0, nt1, G, cta, net, idx;
    __device__ TileSeq(int rows0, int rows1, int pair = 0) {   // pair: 256-row tiles, one worker per cluster of two CTAs
        const int tm = pair ? 2 * BM : BM;
        nt0 = (rows0 + tm - 1) / tm; nt1 = (rows1 + tm - 1) / tm;
        G = (int)gridDim.x >> pair; cta = (int)blockIdx.x >> pair; net = 0; idx = cta;
    }
    __device__ bool next(int& n, int& tile) {
        while (net < 2) {
            if (idx < (net == 0 ? nt0 : nt1)) { n = net; tile = idx; idx += G; return true; }
            ++net;
            idx = G - 1 - cta;
        }
        return false;
    }
};

// hand a 128 x 8 slice (this warp's 32 rows) of the next layer's A operand over: hi -> TMEM in place, lo -> shared-memory ring
__device__ __forceinline__ void publish_slice(const float* v, uint32_t taddr, uint32_t lo_stage, int row_in_tile, int g) {
    uint32_t hi[8];
    float lo[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        const float h = tf32_rna_fast(v[j]);
        hi[j] = __float_as_uint(h);
        lo[j] = tf32_rna_fast(v[j] - h);
    }
    tmem_st8(taddr, hi);
    const uint32_t rbase = lo_stage + (uint32_t)row_in_tile * 128u, sw = (uint32_t)(row_in_tile & 7);
    sts_v4(rbase + ((((uint32_t)(2 * g)) ^ sw) << 4), make_float4(lo[0], lo[1], lo[2], lo[3]));
    sts_v4(rbase + ((((uint32_t)(2 * g + 1)) ^ sw) << 4), make_float4(lo[4], lo[5], lo[6], lo[7]));
}
__device__ __forceinline__ void publish_done(uint64_t* bar, int lane, bool to_leader = false) {
    tmem_wait_st();
    fence_async_smem();                                   // generic-proxy lo writes -> visible to the tensor core's reads
    asm volatile("tcgen05.fence::before_thread_sync;");
    __syncwarp();
    if (lane == 0) { if (to_leader) mbar_arrive_leader(bar); else mbar_arrive(bar); }
}

// =====================================================================================================================
// forward
// =====================================================================================================================
struct alignas(64) FwdNet {
    CUtensorMap mXh, mXl;                 // pre-split input [rows, 64], box 32 x 128
    CUtensorMap mW1h, mW1l;               // [256, 64]   box 32 x 256
    CUtensorMap mW2h, mW2l;               // [n2, 256]   box 32 x n2
    CUtensorMap mW3h, mW3l;               // [128, n2]   box 32 x 128
    const float *b1, *b2, *b3;
    float *H1, *H2, *H3;                  // post-ELU activations [rows, 256], [rows, n2], [rows, 128]
    int rows, n2, exact, pad_;            // exact: expm1f instead of elu_fast
};
struct alignas(64) FwdParams { FwdNet net[2]; };

static constexpr int EPI_W = 16;
static constexpr int F_EPI0 = 2;                           // first epilogue warp
static constexpr int F_THREADS = 32 * (F_EPI0 + EPI_W);    // TMA, MMA, 16 epilogue warps
static constexpr int F_LO_STAGES = 2;
static constexpr int F_X_HI = 0, F_X_LO = 2 * KB_BYTES, F_B = 4 * KB_BYTES, F_LO = F_B + NUNITS * UNIT_BYTES,
                     F_BAR = F_LO + F_LO_STAGES * KB_BYTES, F_SMEM = F_BAR + 256 + 1024;
static_assert(F_SMEM <= 232448, "shared memory budget");

// ELU in 5 straight-line instructions: x > 0 ? x : ex2.approx(x log2 e) - 1.  ABSOLUTE error <= ~3e-7 (2 ulp of ex2.approx around 1
// plus the rounding of the scaled argument) - a few fp32 ulps of the O(1) activations; unlike expm1f it is not small RELATIVE to tiny
// outputs, which no consumer needs: h only enters dot products and ELU' = h + 1.  No branches: a run-time "exact" switch put every
// element into its own basic block (BSSY / BSYNC) and serialised the eight MUFU latencies of a slice (measured: the epilogue's
// tcgen05.ld + bias + ELU part was 55 % of the kernel).
__device__ __forceinline__ float elu_sel(float x, int) {
    float e;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(x * 1.4426950408889634f));
    return (x > 0.0f) ? x : (e - 1.0f);
}

template <int PAIR>
__global__ void __launch_bounds__(F_THREADS, 1) k_mlp_fwd(const __grid_constant__ FwdParams P) {
    // PAIR: every weight k-block is ONE unit per CTA (its half), so three units are as deep as four were; the 32 KB saved double the
    // lo ring: the leader's MMA needs BOTH CTAs' epilogues per k-block, and two stages exposed that hand-shake's latency
    constexpr int NUNITS = PAIR ? 3 : 4, F_LO_STAGES = PAIR ? 4 : 2;
    constexpr int F_LO = F_B + NUNITS * UNIT_BYTES, F_BAR = F_LO + F_LO_STAGES * KB_BYTES;
    static_assert(F_BAR + 256 + 1024 <= F_SMEM, "shared memory layout");
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    uint64_t* x_full = (uint64_t*)(smem + F_BAR);   // (PAIR: the leader's copy collects both CTAs' tiles)
    uint64_t* x_empty = x_full + 1;
    uint64_t* b_full = x_empty + 1;          // [NUNITS]   (PAIR: leader's copy collects both CTAs' halves)
    uint64_t* b_empty = b_full + NUNITS;     // [NUNITS]
    uint64_t* lo_full = b_empty + NUNITS;    // [F_LO_STAGES]   (PAIR: leader's copy collects both CTAs' epilogue warps)
    uint64_t* lo_empty = lo_full + F_LO_STAGES;
    uint64_t* accf = lo_empty + F_LO_STAGES; // [3] layer l's accumulator complete
    uint32_t* tmem_slot = (uint32_t*)(accf + 3);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t sbase = smem_u32(smem);
    const uint32_t rank = PAIR ? cluster_ctarank() : 0u;
    constexpr int TM = PAIR ? 2 * BM : BM;       // rows of a work tile

    if (warp == 0 && lane == 0) {
        mbar_init(x_full, PAIR ? 2 : 1); mbar_init(x_empty, 1);
        for (int s = 0; s < NUNITS; ++s) { mbar_init(&b_full[s], PAIR ? 2 : 1); mbar_init(&b_empty[s], 1); }
        for (int s = 0; s < F_LO_STAGES; ++s) { mbar_init(&lo_full[s], EPI_W * (PAIR ? 2 : 1)); mbar_init(&lo_empty[s], 1); }
        for (int s = 0; s < 3; ++s) mbar_init(&accf[s], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {
        if (PAIR) {
            asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(512u));
            asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;");
        } else {
            asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(512u));
            asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;");
    if (PAIR) cluster_sync_all(); else __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;");
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        // ===== TMA producer: the tile's pre-split input, then the weight k-blocks of the three layers in consumption order =====
        // (PAIR: every CTA loads its own 128 rows of X and its HALF of the rows of each weight k-block: hi at +0, lo at +16 KB of a unit)
        if (lane == 0) {
            uint32_t u = 0, t = 0;
            auto wide = [&](const CUtensorMap* mh, const CUtensorMap* ml, int kb) {   // [256 x 32] hi and lo: one unit each
                for (int half = 0; half < 2; ++half, ++u) {
                    const uint32_t s = u % NUNITS, ph = (u / NUNITS) & 1;
                    mbar_wait_wd(&b_empty[s], ph ^ 1, 100 + (int)s);
                    mbar_expect_tx(&b_full[s], UNIT_BYTES);
                    tma_load_2d(half ? ml : mh, &b_full[s], sbase + F_B + s * UNIT_BYTES, kb * BK, 0);
                }
            };
            auto narrow = [&](const CUtensorMap* mh, const CUtensorMap* ml, int kb, int nrows) { // hi + lo of [nrows x 32] in one unit
                const uint32_t s = u % NUNITS, ph = (u / NUNITS) & 1;
                mbar_wait_wd(&b_empty[s], ph ^ 1, 110 + (int)s);
                const int mine = PAIR ? nrows / 2 : nrows;
                expect_collect<PAIR>(&b_full[s], (uint32_t)(2 * mine * BK * 4), rank);
                tma_collect<PAIR>(mh, &b_full[s], sbase + F_B + s * UNIT_BYTES, kb * BK, (int)rank * mine);
                tma_collect<PAIR>(ml, &b_full[s], sbase + F_B + s * UNIT_BYTES + UNIT_BYTES / 2, kb * BK, (int)rank * mine);
                ++u;
            };
            TileSeq seq(P.net[0].rows, P.net[1].rows, PAIR);
            int ni, tile;
            while (seq.next(ni, tile)) {
                const FwdNet& N = P.net[ni];
                const int m0 = tile * TM + (int)rank * BM;
                mbar_wait_wd(x_empty, (t & 1) ^ 1, 120);
                expect_collect<PAIR>(x_full, 4 * KB_BYTES, rank);
                tma_collect<PAIR>(&N.mXh, x_full, sbase + F_X_HI, 0, m0);
                tma_collect<PAIR>(&N.mXh, x_full, sbase + F_X_HI + KB_BYTES, BK, m0);
                tma_collect<PAIR>(&N.mXl, x_full, sbase + F_X_LO, 0, m0);
                tma_collect<PAIR>(&N.mXl, x_full, sbase + F_X_LO + KB_BYTES, BK, m0);
                for (int kb = 0; kb < 2; ++kb) { if (PAIR) narrow(&N.mW1h, &N.mW1l, kb, 256); else wide(&N.mW1h, &N.mW1l, kb); }
                for (int kb = 0; kb < 8; ++kb) { if (!PAIR && N.n2 == 256) wide(&N.mW2h, &N.mW2l, kb); else narrow(&N.mW2h, &N.mW2l, kb, N.n2); }
                for (int kb = 0; kb < N.n2 / BK; ++kb) narrow(&N.mW3h, &N.mW3l, kb, 128);
                ++t;
            }
        }
    } else if (warp == 1) {
        // ===== MMA issuer (PAIR: the leader CTA issues for both) =====
        if (lane == 0 && rank == 0) {
            uint32_t u = 0, t = 0, li = 0;
            CTL_DECL;
            // weight k-block of an N-row layer: waits for its unit(s), returns the operand addresses and the barriers to release
            auto weights = [&](int nn, uint32_t& b_hi, uint32_t& b_lo, uint64_t*& e0, uint64_t*& e1, int code) {
                e1 = nullptr;
                if (!PAIR && nn == 256) {
                    const uint32_t sa = u % NUNITS, pa = (u / NUNITS) & 1, sb = (u + 1) % NUNITS, pb = ((u + 1) / NUNITS) & 1;
                    u += 2;
                    CTL_WAIT(2, mbar_wait_wd(&b_full[sa], pa, code));
                    CTL_WAIT(2, mbar_wait_wd(&b_full[sb], pb, code + 1));
                    b_hi = sbase + F_B + sa * UNIT_BYTES; b_lo = sbase + F_B + sb * UNIT_BYTES;
                    e0 = &b_empty[sa]; e1 = &b_empty[sb];
                } else {
                    const uint32_t sa = u % NUNITS, pa = (u / NUNITS) & 1;
                    u += 1;
                    CTL_WAIT(2, wait_collect<PAIR>(&b_full[sa], pa, code + 2));
                    b_hi = sbase + F_B + sa * UNIT_BYTES; b_lo = b_hi + UNIT_BYTES / 2;
                    e0 = &b_empty[sa];
                }
            };
            TileSeq seq(P.net[0].rows, P.net[1].rows, PAIR);
            int ni, tile;
            while (seq.next(ni, tile)) {
                const int n2 = P.net[ni].n2;
                const uint32_t par = t & 1;
                const uint32_t c1 = tmem_base + par * 256, c2 = tmem_base + (1 - par) * 256, c3 = c1;
                // ---- layer 1 (A = X from shared memory)
                CTL_WAIT(0, wait_collect<PAIR>(x_full, par, 200));
                if (t > 0) CTL_WAIT(1, mbar_wait_wd(&accf[2], (t - 1) & 1, 201));   // layer 3 of the previous tile has finished READING h2 from c1's half
                asm volatile("tcgen05.fence::after_thread_sync;");
                {
                    // K = 64 is two k-blocks, both resident (X tile and all four weight units): issue the 16 small cross-term MMAs of
                    // BOTH k-blocks first and the 8 dominant a_hi * b_hi MMAs last.  TMEM accumulation truncates on every add, by up to
                    // an ulp of the accumulator: the small terms now land while the accumulator is ~2^-11 of its final size, which
                    // leaves 8 instead of 24 full-size truncations on this layer (the 128-wide layers get the same effect from their
                    // second accumulator).
                    uint32_t b_hi[2], b_lo[2];
                    uint64_t *e0[2], *e1[2];
                    for (int kb = 0; kb < 2; ++kb) weights(256, b_hi[kb], b_lo[kb], e0[kb], e1[kb], 210);
                    asm volatile("tcgen05.fence::after_thread_sync;");
                    constexpr uint32_t id = idesc_tf32_m(TM, 256);
#pragma unroll
                    for (int kb = 0; kb < 2; ++kb) {
                        const uint32_t a_hi = sbase + F_X_HI + kb * KB_BYTES, a_lo = sbase + F_X_LO + kb * KB_BYTES;
#pragma unroll
                        for (int k = 0; k < 4; ++k) {
                            const uint32_t off = k * 32;
                            mma_ss<PAIR>(c1, desc_kmajor(a_lo + off), desc_kmajor(b_hi[kb] + off), id, (kb | k) ? 1u : 0u);
                            mma_ss<PAIR>(c1, desc_kmajor(a_hi + off), desc_kmajor(b_lo[kb] + off), id, 1u);
                        }
                    }
#pragma unroll
                    for (int kb = 0; kb < 2; ++kb) {
                        const uint32_t a_hi = sbase + F_X_HI + kb * KB_BYTES;
#pragma unroll
                        for (int k = 0; k < 4; ++k) mma_ss<PAIR>(c1, desc_kmajor(a_hi + k * 32), desc_kmajor(b_hi[kb] + k * 32), id, 1u);
                    }
                    for (int kb = 0; kb < 2; ++kb) {
                        commit<PAIR>(e0[kb]);
                        if (e1[kb]) commit<PAIR>(e1[kb]);
                    }
                }
                commit<PAIR>(x_empty);
                commit<PAIR>(&accf[0]);
                // ---- layers 2 and 3 (A hi from TMEM in place, A lo from the ring)
                for (int layer = 0; layer < 2; ++layer) {
                    const int nk = layer == 0 ? 8 : n2 / BK, nn = layer == 0 ? n2 : 128;
                    const uint32_t ca = layer == 0 ? c1 : c2, cd = layer == 0 ? c2 : c3;
                    const uint32_t id = idesc_tf32_m(TM, nn);
                    for (int kb = 0; kb < nk; ++kb, ++li) {
                        uint32_t b_hi, b_lo;
                        uint64_t *e0, *e1;
                        weights(nn, b_hi, b_lo, e0, e1, 220);
                        const uint32_t ls = li % F_LO_STAGES, lph = (li / F_LO_STAGES) & 1;
                        CTL_WAIT(3 + layer, wait_collect<PAIR>(&lo_full[ls], lph, 230 + layer));   // the epilogue(s) have published this k-block (hi in TMEM, lo in the ring)
                        asm volatile("tcgen05.fence::after_thread_sync;");
                        const uint32_t a_lo = sbase + F_LO + ls * KB_BYTES;
                        // 128-wide outputs leave 128 free columns next to the accumulator: the dominant a_hi * b_hi products get their own
                        // accumulator (TMEM accumulation truncates: 1/3 of the adds on the large accumulator = 1/3 of the bias), the
                        // two small cross terms share the second one; the epilogue adds them (round to nearest)
                        const uint32_t cs = (nn == 128 && !(P.net[ni].exact & 1)) ? cd + 128 : cd;
#pragma unroll
                        for (int k = 0; k < 4; ++k) {
                            const uint32_t off = k * 32, a_t = ca + kb * BK + k * 8;
                            mma_ss<PAIR>(cs, desc_kmajor(a_lo + off), desc_kmajor(b_hi + off), id, (kb | k) ? 1u : 0u);
                            mma_ts<PAIR>(cs, a_t, desc_kmajor(b_lo + off), id, 1u);
                            mma_ts<PAIR>(cd, a_t, desc_kmajor(b_hi + off), id, (cs != cd && (kb | k) == 0) ? 0u : 1u);
                        }
                        commit<PAIR>(&lo_empty[ls]);
                        commit<PAIR>(e0);
                        if (e1) commit<PAIR>(e1);
                    }
                    commit<PAIR>(&accf[1 + layer]);
                }
                ++t;
            }
            CTL_FLUSH(0, 0, 5);
        }
    } else {
        // ===== epilogue warps: TMEM lane quarter q (rows), 8-column slice g of every 32-column k-block =====
        const int q = warp & 3, g = (warp - F_EPI0) >> 2;
        const uint32_t lane_off = (uint32_t)(q * 32) << 16;
        const int rit = q * 32 + lane;                 // row in this CTA's 128-row tile
        uint32_t t = 0, li = 0;
        CTL_DECL;
        TileSeq seq(P.net[0].rows, P.net[1].rows, PAIR);
        int ni, tile;
        while (seq.next(ni, tile)) {
            const FwdNet& N = P.net[ni];
            const int n2 = N.n2, exact = N.exact;
            const uint32_t par = t & 1;
            const uint32_t c1 = tmem_base + par * 256 + lane_off, c2 = tmem_base + (1 - par) * 256 + lane_off, c3 = c1;
            const int row = tile * TM + (int)rank * BM + rit;
            const bool row_ok = row < N.rows;
            const size_t rsafe = (size_t)(row_ok ? row : 0);
#pragma unroll 1
            for (int layer = 0; layer < 3; ++layer) {
                const int width = layer == 0 ? 256 : (layer == 1 ? n2 : 128);
                const uint32_t cacc = layer == 0 ? c1 : (layer == 1 ? c2 : c3);
                const float* bias = layer == 0 ? N.b1 : (layer == 1 ? N.b2 : N.b3);
                float* out = (layer == 0 ? N.H1 : (layer == 1 ? N.H2 : N.H3)) + rsafe * width;
                CTL_WAIT(layer, mbar_wait_wd(&accf[layer], par, 300 + layer));
                asm volatile("tcgen05.fence::after_thread_sync;");
                // software pipeline over the k-blocks: the accumulator slice and the bias of k-block kb + 1 are requested before kb is
                // processed (tcgen05.ld is asynchronous until tcgen05.wait::ld), so their latency hides behind the ELU / publish of kb
                const bool two_acc = layer > 0 && width == 128 && !(exact & 1);   // big + small accumulator (see the MMA issuer)
                const int nkb = width / BK;
                uint32_t rn[8], rn2[8];
                float4 bn0, bn1;
                tmem_ld8_nowait(cacc + g * 8, rn);
                if (two_acc) tmem_ld8_nowait(cacc + 128 + g * 8, rn2);
                bn0 = __ldg(reinterpret_cast<const float4*>(bias + g * 8)); bn1 = __ldg(reinterpret_cast<const float4*>(bias + g * 8 + 4));
#pragma unroll 1
                for (int kb = 0; kb < nkb; ++kb) {
                    const int col = kb * BK + g * 8;
                    uint32_t r[8];
                    if (two_acc) tmem_wait_ld16(rn, rn2); else tmem_wait_ld8(rn);
#pragma unroll
                    for (int j = 0; j < 8; ++j) r[j] = two_acc ? __float_as_uint(__uint_as_float(rn[j]) + __uint_as_float(rn2[j])) : rn[j];
                    const float4 b0 = bn0, b1 = bn1;
                    if (kb + 1 < nkb) {
                        tmem_ld8_nowait(cacc + col + BK, rn);
                        if (two_acc) tmem_ld8_nowait(cacc + 128 + col + BK, rn2);
                        bn0 = __ldg(reinterpret_cast<const float4*>(bias + col + BK)); bn1 = __ldg(reinterpret_cast<const float4*>(bias + col + BK + 4));
                    }
                    float v[8];
                    v[0] = elu_sel(__uint_as_float(r[0]) + b0.x, exact); v[1] = elu_sel(__uint_as_float(r[1]) + b0.y, exact);
                    v[2] = elu_sel(__uint_as_float(r[2]) + b0.z, exact); v[3] = elu_sel(__uint_as_float(r[3]) + b0.w, exact);
                    v[4] = elu_sel(__uint_as_float(r[4]) + b1.x, exact); v[5] = elu_sel(__uint_as_float(r[5]) + b1.y, exact);
                    v[6] = elu_sel(__uint_as_float(r[6]) + b1.z, exact); v[7] = elu_sel(__uint_as_float(r[7]) + b1.w, exact);
                    if (layer < 2) {
                        const uint32_t ls = li % F_LO_STAGES, lph = (li / F_LO_STAGES) & 1;
                        CTL_WAIT(3, mbar_wait_wd(&lo_empty[ls], lph ^ 1, 310 + layer));   // the MMAs that read this ring stage have completed
                        CTL_WAIT(5, publish_slice(v, cacc + col, sbase + F_LO + ls * KB_BYTES, rit, g));
                        CTL_WAIT(6, publish_done(&lo_full[ls], lane, PAIR && rank != 0));
                        ++li;
                    }
                    // the HBM copy goes out AFTER the hand-over: the fence in publish_done (MEMBAR.ALL.CTA) waits for every earlier
                    // memory operation of the thread, and a global store in front of it put an L2 round trip on the critical path of
                    // every k-block
#ifdef B200_CHAIN_DRY
                    if (exact & 4) { if (v[0] == 123.456f) out[0] = v[1]; } else   // measurement: no global stores
#endif
                    CTL_WAIT(4, if (row_ok) stg_v8(out + col, v));
                }
            }
            ++t;
        }
        if (threadIdx.x == 32 * F_EPI0) CTL_FLUSH(0, 8, 7);
    }
    asm volatile("tcgen05.fence::before_thread_sync;");
    if (PAIR) cluster_sync_all(); else __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;");
    if (warp == 1) {
        if (PAIR) asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u));
        else asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u));
    }
}

// =====================================================================================================================
// backward (input gradients of the hidden layers + bias gradients)
// =====================================================================================================================
struct alignas(64) BwdNet {
    CUtensorMap mZ3, mH2, mH1;            // dz3 [rows,128], h2 [rows,n2], h1 [rows,256]; box 32 x 128 (the epilogue's "aux" tiles)
    CUtensorMap mW3Th, mW3Tl;             // W3^T [n2, 128]   box 32 x n2
    CUtensorMap mW2Th, mW2Tl;             // W2^T [256, n2]   box 32 x 256
    float *DZ2, *DZ1;                     // [rows, n2], [rows, 256]
    float *db2, *db1;                     // bias gradients of layers 2 and 1 (+= column sums of dz2 / dz1)
    int rows, n2;
};
struct alignas(64) BwdParams { BwdNet net[2]; };

static constexpr int B_EPI0 = 3;                           // warp 0 weight TMA, 1 MMA, 2 aux TMA
static constexpr int B_THREADS = 32 * (B_EPI0 + EPI_W);
static constexpr int B_LO_STAGES = 3, B_AUX_STAGES = 3;
static constexpr int B_B = 0, B_LO = NUNITS * UNIT_BYTES, B_AUX = B_LO + B_LO_STAGES * KB_BYTES, B_BAR = B_AUX + B_AUX_STAGES * KB_BYTES,
                     B_SMEM = B_BAR + 256 + 1024;
static_assert(B_SMEM <= 232448, "shared memory budget");

// column sums of this warp's 32 rows x 8 columns -> atomics on dst[0..7]
__device__ __forceinline__ void colsum8(float* v, int lane, float* dst) {
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
    if (lane < 8) atomicAdd(dst + lane, tot);   // lane l holds column l (bit i of l picked the upper half at step 2^i)
}

template <int PAIR>
__global__ void __launch_bounds__(B_THREADS, 1) k_mlp_bwd(const __grid_constant__ BwdParams P) {
    constexpr int NUNITS = PAIR ? 3 : 4, B_LO_STAGES = PAIR ? 4 : 3, B_AUX_STAGES = PAIR ? 4 : 3;   // (see k_mlp_fwd)
    constexpr int B_LO = NUNITS * UNIT_BYTES, B_AUX = B_LO + B_LO_STAGES * KB_BYTES, B_BAR = B_AUX + B_AUX_STAGES * KB_BYTES;
    static_assert(B_BAR + 256 + 1024 <= B_SMEM, "shared memory layout");
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    uint64_t* b_full = (uint64_t*)(smem + B_BAR);   // (PAIR: leader's copy collects both CTAs' halves)
    uint64_t* b_empty = b_full + NUNITS;
    uint64_t* lo_full = b_empty + NUNITS;           // (PAIR: leader's copy collects both CTAs' epilogue warps)
    uint64_t* lo_empty = lo_full + B_LO_STAGES;
    uint64_t* aux_full = lo_empty + B_LO_STAGES;    // (always local: each CTA's epilogue reads its own rows)
    uint64_t* aux_empty = aux_full + B_AUX_STAGES;
    uint64_t* accf = aux_empty + B_AUX_STAGES;   // [2]: dh2 complete, dh1 complete
    uint32_t* tmem_slot = (uint32_t*)(accf + 2);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t sbase = smem_u32(smem);
    const uint32_t rank = PAIR ? cluster_ctarank() : 0u;
    constexpr int TM = PAIR ? 2 * BM : BM;

    if (warp == 0 && lane == 0) {
        for (int s = 0; s < NUNITS; ++s) { mbar_init(&b_full[s], PAIR ? 2 : 1); mbar_init(&b_empty[s], 1); }
        for (int s = 0; s < B_LO_STAGES; ++s) { mbar_init(&lo_full[s], EPI_W * (PAIR ? 2 : 1)); mbar_init(&lo_empty[s], 1); }
        for (int s = 0; s < B_AUX_STAGES; ++s) { mbar_init(&aux_full[s], 1); mbar_init(&aux_empty[s], EPI_W); }
        mbar_init(&accf[0], 1); mbar_init(&accf[1], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {
        if (PAIR) {
            asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(512u));
            asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;");
        } else {
            asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(512u));
            asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;");
    if (PAIR) cluster_sync_all(); else __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;");
    const uint32_t tmem_base = *tmem_slot;
    constexpr uint32_t CZ = 0, CA = 128, CD_LO = 384, CD_HI = 0;   // TMEM columns: dz3, dh2/dz2, dh1 columns [0,128) / [128,256)

    if (warp == 0) {
        // ===== TMA producer: weight k-blocks (W3^T then W2^T per tile) =====
        // single CTA: [256 x 32] tiles take two units (hi, lo), [128 x 32] one (hi at +0, lo at +16 KB).
        // PAIR: every k-block is one unit holding this CTA's HALF of the rows: hi at +0, lo at +16 KB; for W2^T (whose 256 output columns
        //       are produced as two 128-column MMAs) the half is this CTA's 64 rows of each output half: [h0 rows | h1 rows].
        if (lane == 0) {
            uint32_t u = 0;
            auto wide = [&](const CUtensorMap* mh, const CUtensorMap* ml, int kb) {
                for (int half = 0; half < 2; ++half, ++u) {
                    const uint32_t s = u % NUNITS, ph = (u / NUNITS) & 1;
                    mbar_wait_wd(&b_empty[s], ph ^ 1, 500 + (int)s);
                    mbar_expect_tx(&b_full[s], UNIT_BYTES);
                    tma_load_2d(half ? ml : mh, &b_full[s], sbase + B_B + s * UNIT_BYTES, kb * BK, 0);
                }
            };
            auto narrow = [&](const CUtensorMap* mh, const CUtensorMap* ml, int kb, int nrows) {
                const uint32_t s = u % NUNITS, ph = (u / NUNITS) & 1;
                mbar_wait_wd(&b_empty[s], ph ^ 1, 510 + (int)s);
                const int mine = PAIR ? nrows / 2 : nrows;
                expect_collect<PAIR>(&b_full[s], (uint32_t)(2 * mine * BK * 4), rank);
                tma_collect<PAIR>(mh, &b_full[s], sbase + B_B + s * UNIT_BYTES, kb * BK, (int)rank * mine);
                tma_collect<PAIR>(ml, &b_full[s], sbase + B_B + s * UNIT_BYTES + UNIT_BYTES / 2, kb * BK, (int)rank * mine);
                ++u;
            };
            auto halves = [&](const CUtensorMap* mh, const CUtensorMap* ml, int kb) {   // PAIR only: 64-row boxes
                const uint32_t s = u % NUNITS, ph = (u / NUNITS) & 1;
                mbar_wait_wd(&b_empty[s], ph ^ 1, 515 + (int)s);
                expect_collect<PAIR>(&b_full[s], (uint32_t)UNIT_BYTES, rank);
                const uint32_t base = sbase + B_B + s * UNIT_BYTES;
                for (int h = 0; h < 2; ++h) {
                    tma_collect<PAIR>(mh, &b_full[s], base + h * (UNIT_BYTES / 4), kb * BK, h * 128 + (int)rank * 64);
                    tma_collect<PAIR>(ml, &b_full[s], base + UNIT_BYTES / 2 + h * (UNIT_BYTES / 4), kb * BK, h * 128 + (int)rank * 64);
                }
                ++u;
            };
            TileSeq seq(P.net[0].rows, P.net[1].rows, PAIR);
            int ni, tile;
            while (seq.next(ni, tile)) {
                const BwdNet& N = P.net[ni];
                for (int kb = 0; kb < 4; ++kb) { if (!PAIR && N.n2 == 256) wide(&N.mW3Th, &N.mW3Tl, kb); else narrow(&N.mW3Th, &N.mW3Tl, kb, N.n2); }
                for (int kb = 0; kb < N.n2 / BK; ++kb) { if (PAIR) halves(&N.mW2Th, &N.mW2Tl, kb); else wide(&N.mW2Th, &N.mW2Tl, kb); }
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
            TileSeq seq(P.net[0].rows, P.net[1].rows, PAIR);
            int cn, ct, nn = 0, nt = 0;
            bool have = seq.next(cn, ct);
            if (have) for (int kb = 0; kb < 4; ++kb) aux(&P.net[cn].mZ3, kb, ct * TM + (int)rank * BM);
            while (have) {
                const bool hn = seq.next(nn, nt);
                const BwdNet& N = P.net[cn];
                const int m0 = ct * TM + (int)rank * BM;
                for (int kb = 0; kb < N.n2 / BK; ++kb) aux(&N.mH2, kb, m0);
                for (int kb = 4; kb < 8; ++kb) aux(&N.mH1, kb, m0);
                if (hn) for (int kb = 0; kb < 4; ++kb) aux(&P.net[nn].mZ3, kb, nt * TM + (int)rank * BM);
                for (int kb = 0; kb < 4; ++kb) aux(&N.mH1, kb, m0);
                have = hn; cn = nn; ct = nt;
            }
        }
    } else if (warp == 1) {
        // ===== MMA issuer (PAIR: the leader CTA issues for both) =====
        if (lane == 0 && rank == 0) {
            uint32_t u = 0, li = 0;
            CTL_DECL;
            TileSeq seq(P.net[0].rows, P.net[1].rows, PAIR);
            int ni, tile;
            while (seq.next(ni, tile)) {
                const int n2 = P.net[ni].n2;
                // ---- dh2 [128, n2] = dz3 [128,128] W3: 4 k-blocks
                for (int kb = 0; kb < 4; ++kb, ++li) {
                    uint32_t b_hi, b_lo;
                    uint64_t *e0, *e1 = nullptr;
                    if (!PAIR && n2 == 256) {
                        const uint32_t sa = u % NUNITS, pa = (u / NUNITS) & 1, sb = (u + 1) % NUNITS, pb = ((u + 1) / NUNITS) & 1;
                        u += 2;
                        CTL_WAIT(2, mbar_wait_wd(&b_full[sa], pa, 600));
                        CTL_WAIT(2, mbar_wait_wd(&b_full[sb], pb, 601));
                        b_hi = sbase + B_B + sa * UNIT_BYTES; b_lo = sbase + B_B + sb * UNIT_BYTES;
                        e0 = &b_empty[sa]; e1 = &b_empty[sb];
                    } else {
                        const uint32_t sa = u % NUNITS, pa = (u / NUNITS) & 1;
                        u += 1;
                        CTL_WAIT(2, wait_collect<PAIR>(&b_full[sa], pa, 602));
                        b_hi = sbase + B_B + sa * UNIT_BYTES; b_lo = b_hi + UNIT_BYTES / 2;
                        e0 = &b_empty[sa];
                    }
                    const uint32_t ls = li % B_LO_STAGES, lph = (li / B_LO_STAGES) & 1;
                    CTL_WAIT(3, wait_collect<PAIR>(&lo_full[ls], lph, 610));
                    asm volatile("tcgen05.fence::after_thread_sync;");
                    const uint32_t a_lo = sbase + B_LO + ls * KB_BYTES;
                    const uint32_t id = idesc_tf32_m(TM, n2);
#pragma unroll
                    for (int k = 0; k < 4; ++k) {
                        const uint32_t off = k * 32, a_t = tmem_base + CZ + kb * BK + k * 8;
                        mma_ss<PAIR>(tmem_base + CA, desc_kmajor(a_lo + off), desc_kmajor(b_hi + off), id, (kb | k) ? 1u : 0u);
                        mma_ts<PAIR>(tmem_base + CA, a_t, desc_kmajor(b_lo + off), id, 1u);
                        mma_ts<PAIR>(tmem_base + CA, a_t, desc_kmajor(b_hi + off), id, 1u);
                    }
                    commit<PAIR>(&lo_empty[ls]);
                    commit<PAIR>(e0);
                    if (e1) commit<PAIR>(e1);
                }
                commit<PAIR>(&accf[0]);
                // ---- dh1 [128, 256] = dz2 [128, n2] W2: n2 / 32 k-blocks, the output as two 128-column halves
                for (int kb = 0; kb < n2 / BK; ++kb, ++li) {
                    uint32_t b_hi, b_lo, hstep;
                    uint64_t *e0, *e1 = nullptr;
                    if (PAIR) {
                        const uint32_t sa = u % NUNITS, pa = (u / NUNITS) & 1;
                        u += 1;
                        CTL_WAIT(2, wait_collect<PAIR>(&b_full[sa], pa, 620));
                        b_hi = sbase + B_B + sa * UNIT_BYTES; b_lo = b_hi + UNIT_BYTES / 2; hstep = UNIT_BYTES / 4;   // 64 rows per output half
                        e0 = &b_empty[sa];
                    } else {
                        const uint32_t sa = u % NUNITS, pa = (u / NUNITS) & 1, sb = (u + 1) % NUNITS, pb = ((u + 1) / NUNITS) & 1;
                        u += 2;
                        CTL_WAIT(2, mbar_wait_wd(&b_full[sa], pa, 620));
                        CTL_WAIT(2, mbar_wait_wd(&b_full[sb], pb, 621));
                        b_hi = sbase + B_B + sa * UNIT_BYTES; b_lo = sbase + B_B + sb * UNIT_BYTES; hstep = UNIT_BYTES / 2;   // rows 128..255 of the tile
                        e0 = &b_empty[sa]; e1 = &b_empty[sb];
                    }
                    const uint32_t ls = li % B_LO_STAGES, lph = (li / B_LO_STAGES) & 1;
                    CTL_WAIT(4, wait_collect<PAIR>(&lo_full[ls], lph, 630));
                    asm volatile("tcgen05.fence::after_thread_sync;");
                    const uint32_t a_lo = sbase + B_LO + ls * KB_BYTES;
                    constexpr uint32_t id = idesc_tf32_m(TM, 128);
#pragma unroll
                    for (int k = 0; k < 4; ++k) {
                        const uint32_t off = k * 32, a_t = tmem_base + CA + kb * BK + k * 8;
#pragma unroll
                        for (int h = 0; h < 2; ++h) {
                            const uint32_t cd = tmem_base + (h ? CD_HI : CD_LO), boff = off + h * hstep;
                            mma_ss<PAIR>(cd, desc_kmajor(a_lo + off), desc_kmajor(b_hi + boff), id, (kb | k) ? 1u : 0u);
                            mma_ts<PAIR>(cd, a_t, desc_kmajor(b_lo + boff), id, 1u);
                            mma_ts<PAIR>(cd, a_t, desc_kmajor(b_hi + boff), id, 1u);
                        }
                    }
                    commit<PAIR>(&lo_empty[ls]);
                    commit<PAIR>(e0);
                    if (e1) commit<PAIR>(e1);
                }
                commit<PAIR>(&accf[1]);
            }
            CTL_FLUSH(1, 0, 5);
        }
    } else if (warp >= B_EPI0) {
        // ===== epilogue warps =====
        const int q = warp & 3, g = (warp - B_EPI0) >> 2;
        const uint32_t lane_off = (uint32_t)(q * 32) << 16;
        const int rit = q * 32 + lane;
        const uint32_t sw = (uint32_t)(rit & 7);
        const bool to_leader = PAIR && rank != 0;
        uint32_t t = 0, li = 0, ai = 0;
        CTL_DECL;
        // this thread's 8 floats of the aux k-block in ring stage s (128-byte swizzled rows).  The stage is handed back to the TMA
        // producer by aux_release() only AFTER instructions that consume the loaded registers have issued: an mbarrier arrive right
        // behind the ld.shared can overtake the loads still in flight (measured: 0.15 % of the rows read the next k-block's data).
        auto aux_read = [&](float* h) -> uint32_t {
            const uint32_t s = ai % B_AUX_STAGES, ph = (ai / B_AUX_STAGES) & 1;
            CTL_WAIT(2, mbar_wait_wd(&aux_full[s], ph, 700));
            const uint32_t rbase = sbase + B_AUX + s * KB_BYTES + (uint32_t)rit * 128u;
            const float4 x0 = lds_v4(rbase + ((((uint32_t)(2 * g)) ^ sw) << 4)), x1 = lds_v4(rbase + ((((uint32_t)(2 * g + 1)) ^ sw) << 4));
            h[0] = x0.x; h[1] = x0.y; h[2] = x0.z; h[3] = x0.w; h[4] = x1.x; h[5] = x1.y; h[6] = x1.z; h[7] = x1.w;
            ++ai;
            return s;
        };
        auto aux_release = [&](uint32_t s) {
            __syncwarp();
            if (lane == 0) mbar_arrive(&aux_empty[s]);
        };
        auto publish = [&](const float* v, uint32_t tcol) {
            const uint32_t ls = li % B_LO_STAGES, lph = (li / B_LO_STAGES) & 1;
            CTL_WAIT(3, mbar_wait_wd(&lo_empty[ls], lph ^ 1, 710));
            publish_slice(v, tmem_base + lane_off + tcol, sbase + B_LO + ls * KB_BYTES, rit, g);
            publish_done(&lo_full[ls], lane, to_leader);
            ++li;
        };
        auto stage_dz3 = [&]() {   // dz3 of a tile: aux tiles -> A operand (no arithmetic)
#pragma unroll 1
            for (int kb = 0; kb < 4; ++kb) {
                float v[8];
                const uint32_t as = aux_read(v);
                publish(v, CZ + kb * BK + g * 8);   // (tcgen05.st / st.shared of the loaded values, tcgen05.wait::st)
                aux_release(as);
            }
        };
        // one k-block of an ELU'-masked gradient: v = acc * ELU'(h); store; column sums; optionally hand over to the next GEMM
        auto grad_block = [&](uint32_t tcol, float* out_row, int col, float* db, bool row_ok, bool handoff) {
            float h[8], v[8];
            const uint32_t as = aux_read(h);
            uint32_t r[8];
            tmem_ld8(tmem_base + lane_off + tcol, r);
#pragma unroll
            for (int j = 0; j < 8; ++j) v[j] = __uint_as_float(r[j]) * ((h[j] > 0.0f) ? 1.0f : (h[j] + 1.0f));
            if (handoff) publish(v, tcol);
            if (row_ok) stg_v8(out_row + col, v);   // after the hand-over (see k_mlp_fwd): keeps the L2 round trip off the fence
            if (!row_ok) {
#pragma unroll
                for (int j = 0; j < 8; ++j) v[j] = 0.0f;
            }
            colsum8(v, lane, db + col);
            aux_release(as);
        };
        TileSeq seq(P.net[0].rows, P.net[1].rows, PAIR);
        int cn, ct, nn = 0, nt = 0;
        bool have = seq.next(cn, ct);
        if (have) stage_dz3();
        while (have) {
            const bool hn = seq.next(nn, nt);
            const BwdNet& N = P.net[cn];
            const int n2 = N.n2;
            const int row = ct * TM + (int)rank * BM + rit;
            const bool row_ok = row < N.rows;
            const size_t rsafe = (size_t)(row_ok ? row : 0);
            const uint32_t par = t & 1;
            // dz2
            CTL_WAIT(0, mbar_wait_wd(&accf[0], par, 720));
            asm volatile("tcgen05.fence::after_thread_sync;");
#pragma unroll 1
            for (int kb = 0; kb < n2 / BK; ++kb) grad_block(CA + kb * BK + g * 8, N.DZ2 + rsafe * n2, kb * BK + g * 8, N.db2, row_ok, true);
            // dz1: high half of the columns first (it shares TMEM columns with the next tile's dz3)
            CTL_WAIT(1, mbar_wait_wd(&accf[1], par, 721));
            asm volatile("tcgen05.fence::after_thread_sync;");
#pragma unroll 1
            for (int kb = 4; kb < 8; ++kb) grad_block(CD_HI + (kb - 4) * BK + g * 8, N.DZ1 + rsafe * 256, kb * BK + g * 8, N.db1, row_ok, false);
            if (hn) stage_dz3();
#pragma unroll 1
            for (int kb = 0; kb < 4; ++kb) grad_block(CD_LO + kb * BK + g * 8, N.DZ1 + rsafe * 256, kb * BK + g * 8, N.db1, row_ok, false);
            have = hn; cn = nn; ct = nt;
            ++t;
        }
        if (threadIdx.x == 32 * B_EPI0) CTL_FLUSH(1, 8, 4);
    }
    asm volatile("tcgen05.fence::before_thread_sync;");
    if (PAIR) cluster_sync_all(); else __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;");
    if (warp == 1) {
        if (PAIR) asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u));
        else asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u));
    }
}

}  // namespace chain
}  // namespace b200
