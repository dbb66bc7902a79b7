// gemm_tc.cuh - Blackwell-native fp32-accurate GEMMs for the actor/critic MLP (utils/model.py:9-26): tcgen05.mma
// (kind::tf32) with TMEM accumulators, TMA-staged operands (cp.async.bulk.tensor, 128-byte swizzle), mbarrier pipelines.
//
// Precision: every product is evaluated as  a_lo*b_hi + a_hi*b_lo + a_hi*b_hi  with both halves TF32 - three tcgen05.mma
// per k-step into the same TMEM accumulator (3xTF32; SASS: UTCHMMA, UTMALDG, LDTM).  Activations and their gradients
// live in HBM as plain fp32 (4 B / element); the split happens IN SHARED MEMORY: TMA lands the raw fp32 tile, which the
// tensor core reads directly as the hi half (kind::tf32 ignores the low 13 mantissa bits: measured, truncation), and
// four converter warps write lo = tf32(x - trunc13(x)) into a second tile of the same (swizzled) layout, publish it with
// fence.proxy.async and an mbarrier.  Weights are pre-split (hi rounded to nearest, so the dropped lo*lo term has no
// systematic sign) by k_weight_prep: they are small and shared by all CTAs.
//
// Two kernels:
//   k_tc_rowmajor<BN, STAGES, EPI, NACC>   C[M, Nout] = epi(A[M, K] * B[Nout, K]^T)   both operands K-major
//        persistent CTAs (one per SM, 576 threads), warp roles: 0 = TMA producer, 1 = MMA issuer (single elected thread),
//        2..17 = epilogue.  NACC TMEM accumulators per tile (k-blocks rotate over them and the epilogue adds them with
//        round-to-nearest FADDs: TMEM accumulation truncates), double-buffered across tiles when 2*BN*NACC <= 512 columns
//        so the epilogue of tile i overlaps the main loop of tile i+1.  EPI_FWD: +bias, ELU, store fp32.  EPI_DGRAD:
//        * ELU'(h) with h = the layer's stored post-activation, store fp32, and the bias gradient (column sums) of the
//        layer below.  Global IO is 256-bit per thread (one 32-byte sector).  Warps 18..21 are the lo converters.
//   k_tc_wgrad<BN, STAGES>           dW[Nout, Kin] += dY[m, Nout]^T X[m, Kin] over a contiguous range of rows m per CTA
//        the reduction index is the ROW index of both operands -> MN-major operands; for 32-bit elements the only legal
//        MN-major shared-memory layout is SWIZZLE_128B_BASE32B (TMA: CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B).  Both operands
//        are activations here, so eight converter warps split both; dY is additionally rounded in place (hi = rna(x)) so
//        that the dropped lo*lo term of the two truncated operands does not accumulate a signed bias over 98k rows.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "common.cuh"

#include <map>
#include <tuple>

namespace b200 {
namespace tc {

// Optional per-CTA timeline (-DB200_TC_TIMELINE, tools/tc_timeline.py): globaltimer stamps / accumulated waits per role, written to
// a device array that b200_tc_timeline_read copies out.  Compiled out of the product build.
#ifdef B200_TC_TIMELINE
#define TL_SLOTS 40
__device__ unsigned long long g_tl[TL_SLOTS][160][12];
__device__ __forceinline__ unsigned long long tl_now() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
    return t;
}
#define TL_SET(kind, slot, v) g_tl[g.tl_slot][blockIdx.x + gridDim.x * (blockIdx.y + gridDim.y * blockIdx.z) < 160 ? blockIdx.x + gridDim.x * (blockIdx.y + gridDim.y * blockIdx.z) : 159][slot] = (v)
#define TL_DECL(x) unsigned long long x = 0
#define TL_T0(x) const unsigned long long x = tl_now()
#define TL_ACC(acc, t0) acc += tl_now() - (t0)
#else
#define TL_SET(kind, slot, v)
#define TL_DECL(x)
#define TL_T0(x)
#define TL_ACC(acc, t0)
#endif

static constexpr int BM = 128;  // rows of the output tile = TMEM lanes = UMMA M
static constexpr int BK = 32;   // fp32 elements per k-block = one 128-byte swizzle row

enum { EPI_FWD = 0, EPI_DGRAD = 1 };

struct RowArgs {
    float* out;          // [M, ldo] fp32
    const float* bias;   // EPI_FWD
    const float* aux;    // EPI_DGRAD: post-activation (fp32) of the layer whose input gradient is produced
    float* colsum;       // nullable (EPI_DGRAD): colsum[c] += sum over rows of the stored output column c = the bias gradient of the
                         // layer this dX feeds (utils/runner.py:163 autograd of nn.Linear.bias)
    int M, Nout, K;      // rows, output columns (multiple of 32), reduction length (TMA zero-fills beyond the tensor)
    int ldo;             // leading dimension of out and aux (floats)
    int tl_slot;         // timeline slot (debug builds only)
};

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* b, uint32_t cnt) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(b)), "r"(cnt));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* b, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(b)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* b) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(b)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* b, uint32_t parity) {
    uint32_t done;
    const uint32_t a = smem_u32(b);
    do {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(done) : "r"(a), "r"(parity) : "memory");
    } while (!done);
}
__device__ __forceinline__ void tma_load_2d(const CUtensorMap* map, uint64_t* bar, uint32_t dst, int c0, int c1) {
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
                 ::"r"(dst), "l"((uint64_t)map), "r"(smem_u32(bar)), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void tma_load_3d(const CUtensorMap* map, uint64_t* bar, uint32_t dst, int c0, int c1, int c2) {
    asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
                 ::"r"(dst), "l"((uint64_t)map), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2) : "memory");
}
// shared-memory matrix descriptors (cute::UMMA::SmemDescriptor bit layout; version 1 = Blackwell)
__device__ __forceinline__ uint64_t desc_kmajor(uint32_t saddr) {   // K-major, SWIZZLE_128B: 8-row groups 1024 B apart
    return (uint64_t)((saddr >> 4) & 0x3FFF) | ((uint64_t)1 << 16) | ((uint64_t)(1024 >> 4) << 32) | ((uint64_t)1 << 46) | ((uint64_t)2 << 61);
}
// MN-major, SWIZZLE_128B_BASE32B: 32-float column chunks `lbo` bytes apart (= one TMA box of WBK rows x 128 B), 4-row k groups 512 B apart
__device__ __forceinline__ uint64_t desc_mnmajor(uint32_t saddr, uint32_t lbo) {
    return (uint64_t)((saddr >> 4) & 0x3FFF) | ((uint64_t)(lbo >> 4) << 16) | ((uint64_t)(512 >> 4) << 32) | ((uint64_t)1 << 46) | ((uint64_t)1 << 61);
}
// instruction descriptor (cute::UMMA::InstrDescriptor): D = f32, A = B = tf32, dense, M = 128
__host__ __device__ constexpr uint32_t idesc_tf32(int n, bool mn_major) {
    return (1u << 4) | (2u << 7) | (2u << 10) | (mn_major ? ((1u << 15) | (1u << 16)) : 0u) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);
}
__device__ __forceinline__ void umma_tf32(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(tmem_d), "l"(da), "l"(db), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t* r) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
          "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]),
          "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]),
          "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t* r) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
          "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ float tf32_rna(float x) {
    uint32_t r;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
    return __uint_as_float(r);
}

// lo half of the in-smem split: the tensor core truncates the raw fp32 word to tf32, so lo = x - trunc13(x) (exact, <= 13
// significant bits), rounded to tf32
__device__ __forceinline__ float tf32_rna_fast(float x) {   // round-half-away on the bit pattern (finite inputs): 2 ALU ops instead of cvt.rna's ~6
    return __uint_as_float((__float_as_uint(x) + 0x1000u) & 0xFFFFE000u);
}
__device__ __forceinline__ float lo_trunc(float x) {
    return tf32_rna_fast(x - __uint_as_float(__float_as_uint(x) & 0xFFFFE000u));
}
__device__ __forceinline__ float4 lo_trunc4(const float4 x) {
    return make_float4(lo_trunc(x.x), lo_trunc(x.y), lo_trunc(x.z), lo_trunc(x.w));
}
__device__ __forceinline__ float4 lds_v4(uint32_t a) {
    float4 v;
    asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(a));
    return v;
}
__device__ __forceinline__ void sts_v4(uint32_t a, const float4 v) {
    asm volatile("st.shared.v4.f32 [%0], {%1,%2,%3,%4};" ::"r"(a), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

__device__ __forceinline__ void ldg_v8(const float* p, float* a) {  // 256-bit global load (sm_100: LDG.E.256), 32-byte aligned
    asm volatile("ld.global.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=f"(a[0]), "=f"(a[1]), "=f"(a[2]), "=f"(a[3]), "=f"(a[4]), "=f"(a[5]), "=f"(a[6]), "=f"(a[7]) : "l"(p));
}
__device__ __forceinline__ void stg_v8(float* p, const float* a) {  // 256-bit global store: one full 32-byte sector per thread
    asm volatile("st.global.v8.f32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};"
                 ::"l"(p), "f"(a[0]), "f"(a[1]), "f"(a[2]), "f"(a[3]), "f"(a[4]), "f"(a[5]), "f"(a[6]), "f"(a[7]) : "memory");
}
// ELU for the tcgen05 forward epilogue (critic): x > 0 ? x : expm1(x).  expm1 for x <= 0 as a degree-7 Taylor polynomial on
// [-0.25, 0] (truncation < 2e-9 relative) and ex2.approx(x * log2 e) - 1 below (|result| >= 0.22, absolute error of
// ex2.approx <= 2^-22.5 * e^x -> < 1e-6 relative): branch-free, ~12 instructions instead of ~30 for expm1f.
__device__ __forceinline__ float elu_fast(float x) {
    const float xm = fminf(x, 0.0f);
    float p = 1.0f / 5040.0f;
    p = fmaf(p, xm, 1.0f / 720.0f);
    p = fmaf(p, xm, 1.0f / 120.0f);
    p = fmaf(p, xm, 1.0f / 24.0f);
    p = fmaf(p, xm, 1.0f / 6.0f);
    p = fmaf(p, xm, 0.5f);
    p = fmaf(p, xm, 1.0f);
    p = p * xm;
    float e;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(xm * 1.4426950408889634f));
    const float neg = (xm > -0.25f) ? p : (e - 1.0f);
    return (x > 0.0f) ? x : neg;
}

static constexpr int EPI_WARPS = 16;                   // 4 per TMEM lane quarter, each takes a quarter of the tile's columns
static constexpr int CONV_WARPS = 4;                   // in-smem lo converters of the A (activation) tile
static constexpr int CONV_T0 = 64 + 32 * EPI_WARPS;    // first converter thread
static constexpr int TMAB_T0 = CONV_T0 + 32 * CONV_WARPS;       // the warp that streams the weight (B) tiles
static constexpr int ROW_THREADS = TMAB_T0 + 32;  // TMA-A warp + MMA warp + epilogue warps + converter warps + TMA-B warp
// ---- CTA-pair helpers (cta_group::2: two CTAs of a cluster on the two SMs of a TPC share one 256-row MMA) --------------------
static constexpr uint32_t PEER_MASK = 0xFEFFFFFFu;   // clears the pair-rank bit of a shared::cluster address -> the leader (even) CTA
__device__ __forceinline__ uint32_t cluster_ctarank() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
    asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive_leader(uint64_t* b) {   // arrive on this barrier's twin in the pair's leader CTA
    asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(smem_u32(b) & PEER_MASK) : "memory");
}
__device__ __forceinline__ void mbar_wait_cluster(uint64_t* b, uint32_t parity) {   // waits for arrivals that may come from the peer CTA
    uint32_t done;
    const uint32_t a = smem_u32(b);
    do {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(done) : "r"(a), "r"(parity) : "memory");
    } while (!done);
}
// the load lands in THIS CTA's shared memory, its bytes are counted on the leader CTA's mbarrier
__device__ __forceinline__ void tma_load_2d_pair(const CUtensorMap* map, uint64_t* bar, uint32_t dst, int c0, int c1) {
    asm volatile("cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
                 ::"r"(dst), "l"((uint64_t)map), "r"(smem_u32(bar) & PEER_MASK), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void umma_tf32_pair(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::2.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(tmem_d), "l"(da), "l"(db), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void umma_commit_pair(uint64_t* bar) {   // arrives on `bar` in BOTH CTAs of the pair
    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
                 ::"r"(smem_u32(bar)), "h"((uint16_t)3) : "memory");
}
__host__ __device__ constexpr uint32_t idesc_tf32_m(int m, int n) {   // K-major operands, D = f32, A = B = tf32
    return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(m >> 4) << 24);
}

// PAIR = 1: a cluster of two CTAs works on a 256-row tile (cta_group::2, UMMA M = 256).  Each CTA stages its own 128 rows of A
// and HALF of the B (weight) tile; the tensor cores of the two SMs exchange the B halves, so the per-CTA L2 -> shared-memory fill
// per k-block drops from 16 + 2*32 KB to 16 + 2*16 KB (BN = 256) - the L2 slice throughput (~42 B/clk/SM) was the bound - and
// a stage shrinks from 96 to 64 KB (3 stages instead of 2).  The leader CTA issues the MMAs; mbarriers that collect work of
// both CTAs (B landed, A lo converted, accumulator drained) live in the leader and are signalled remotely by the peer.
// Two rings: the raw activation tiles (from HBM: latency ~1.5 us under load, so NA = 4..8 stages of 16 KB keep 50-100 KB in
// flight per SM) and the weight tiles + converted A lo tiles (from L2 / produced on chip: NB = 2..3 stages).
template <int BN, int NA, int NB, int PAIR>
struct RowSmem {
    static constexpr int A_BYTES = BM * BK * 4, B_BYTES = (BN / (PAIR ? 2 : 1)) * BK * 4, BSTAGE_BYTES = A_BYTES + 2 * B_BYTES;
    static constexpr int B_RING = NA * A_BYTES;                               // offset of the second ring
    static constexpr int BAR = B_RING + NB * BSTAGE_BYTES;                   // offset of the mbarriers
    static constexpr int TOTAL = BAR + 256 + 1024;                            // + barriers + alignment slack
    // A ring stage: [A raw fp32 = hi operand];  B ring stage: [A lo (converter)][B hi][B lo]   (B = this CTA's half of the rows when PAIR)
};

// The epilogue is instruction-bound (ELU, address math), so it gets 16 warps; each thread owns one output row (its TMEM
// lane) and 32 consecutive columns per step, which it reads / writes as 32-byte vectors: one full sector per instruction, no
// shared-memory staging is needed.
template <int BN, int NA, int NB, int EPI, int NACC, int PAIR>
__global__ void __launch_bounds__(ROW_THREADS, 1)
k_tc_rowmajor(const __grid_constant__ CUtensorMap mA, const __grid_constant__ CUtensorMap mBh, const __grid_constant__ CUtensorMap mBl,
              const RowArgs g) {
    using S = RowSmem<BN, NA, NB, PAIR>;
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    uint64_t* afull = (uint64_t*)(smem + S::BAR);   // [NA] this CTA's raw A tile landed
    uint64_t* aempty = afull + NA;      // [NA] the MMAs that read the raw A stage have completed
    uint64_t* bfull = aempty + NA;      // [NB] B hi / lo landed                         (PAIR: leader's copy counts both CTAs' halves)
    uint64_t* bempty = bfull + NB;      // [NB] the MMAs that read the B stage (and its A lo tile) have completed
    uint64_t* conv = bempty + NB;       // [NB] A lo tile(s) written                      (PAIR: leader's copy collects both CTAs)
    uint64_t* tfull = conv + NB;        // [2]
    uint64_t* tempty = tfull + 2;       // [2]                                          (PAIR: leader's copy collects both CTAs)
    uint32_t* tmem_slot = (uint32_t*)(tempty + 2);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t rank = PAIR ? cluster_ctarank() : 0u;
    constexpr int TM = PAIR ? 2 * BM : BM;                       // rows of a work tile
    const int m_tiles = (g.M + TM - 1) / TM, n_tiles = (g.Nout + BN - 1) / BN, tiles = m_tiles * n_tiles;
    const int worker = PAIR ? (int)(blockIdx.x >> 1) : (int)blockIdx.x, workers = PAIR ? (int)(gridDim.x >> 1) : (int)gridDim.x;
    const int nk = (g.K + BK - 1) / BK;
    // NACC = 2 ("accurate"): TMEM accumulation truncates, so the dominant a_hi*b_hi products get their own accumulator (chain
    // length K/8) and the two small cross terms a second one; the epilogue adds them with a round-to-nearest FADD.  One MMA of
    // N = 2*BN over the contiguous [B hi; B lo] tile produces a_hi*b_hi | a_hi*b_lo side by side (A hi is read once), a second
    // MMA of N = BN adds a_lo*b_hi to the small half.  Accumulators are double-buffered across tiles when TMEM has room.
    static_assert(NACC == 1 || (NACC == 2 && !PAIR && 2 * BN <= 256), "accumulator layout");
    constexpr int NBUF = (2 * BN * NACC <= 512) ? 2 : 1;
    constexpr uint32_t TCOLS = NBUF * BN * NACC;
    static_assert(TCOLS <= 512 && (TCOLS & (TCOLS - 1)) == 0, "TMEM allocation must be a power of two <= 512 columns");

    if (warp == 0 && lane == 0) {
        for (int s = 0; s < NA; ++s) { mbar_init(&afull[s], 1); mbar_init(&aempty[s], 1); }
        for (int s = 0; s < NB; ++s) {
            mbar_init(&bfull[s], PAIR ? 2 : 1); mbar_init(&bempty[s], 1); mbar_init(&conv[s], 32 * CONV_WARPS * (PAIR ? 2 : 1));
        }
        for (int a = 0; a < 2; ++a) { mbar_init(&tfull[a], 1); mbar_init(&tempty[a], EPI_WARPS * (PAIR ? 2 : 1)); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {
        if (PAIR) {
            asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(TCOLS));
            asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;");
        } else {
            asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(TCOLS));
            asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;");
    if (PAIR) cluster_sync_all(); else __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;");
    const uint32_t tmem_base = *tmem_slot;
    if (threadIdx.x == 0) TL_SET(0, 0, tl_now());

    if (warp == 0) {
        // ===== TMA producer A: raw activation tiles (HBM) =====
        if (lane == 0) {
            TL_DECL(w_empty);
            uint32_t it = 0;
            for (int tile = worker; tile < tiles; tile += workers) {
                const int m0 = (tile / n_tiles) * TM + (int)rank * BM;
                for (int kb = 0; kb < nk; ++kb, ++it) {
                    const uint32_t s = it % NA, ph = (it / NA) & 1;
                    TL_T0(t0);
                    mbar_wait(&aempty[s], ph ^ 1);
                    TL_ACC(w_empty, t0);
                    mbar_expect_tx(&afull[s], S::A_BYTES);
                    tma_load_2d(&mA, &afull[s], smem_u32(smem + s * S::A_BYTES), kb * BK, m0);
                }
            }
            TL_SET(0, 1, tl_now());
            TL_SET(0, 2, w_empty);
        }
    } else if (warp == TMAB_T0 / 32) {
        // ===== TMA producer B: pre-split weight tiles (L2) =====
        if (lane == 0) {
            uint32_t it = 0;
            for (int tile = worker; tile < tiles; tile += workers) {
                const int n0 = (tile % n_tiles) * BN + (int)rank * (BN / 2) * PAIR;
                for (int kb = 0; kb < nk; ++kb, ++it) {
                    const uint32_t s = it % NB, ph = (it / NB) & 1;
                    mbar_wait(&bempty[s], ph ^ 1);
                    const uint32_t st = smem_u32(smem + S::B_RING + s * S::BSTAGE_BYTES) + S::A_BYTES;
                    if (PAIR) {
                        if (rank == 0) mbar_expect_tx(&bfull[s], 4 * S::B_BYTES);   // hi + lo halves of both CTAs
                        else mbar_arrive_leader(&bfull[s]);
                        tma_load_2d_pair(&mBh, &bfull[s], st, kb * BK, n0);
                        tma_load_2d_pair(&mBl, &bfull[s], st + S::B_BYTES, kb * BK, n0);
                    } else {
                        mbar_expect_tx(&bfull[s], 2 * S::B_BYTES);
                        tma_load_2d(&mBh, &bfull[s], st, kb * BK, n0);
                        tma_load_2d(&mBl, &bfull[s], st + S::B_BYTES, kb * BK, n0);
                    }
                }
            }
        }
    } else if (warp == 1) {
        // ===== MMA issuer (PAIR: the leader CTA issues for both) =====
        if (lane == 0 && rank == 0) {
            constexpr uint32_t idesc = idesc_tf32_m(TM, BN);
            TL_DECL(w_tempty); TL_DECL(w_full); TL_DECL(w_conv);
            uint32_t it = 0, tl = 0;
            for (int tile = worker; tile < tiles; tile += workers, ++tl) {
                const uint32_t a = (NBUF == 2) ? (tl & 1) : 0, aph = (NBUF == 2) ? ((tl >> 1) & 1) : (tl & 1);
                TL_T0(t0);
                if (PAIR) mbar_wait_cluster(&tempty[a], aph ^ 1); else mbar_wait(&tempty[a], aph ^ 1);  // the epilogue has drained this accumulator set
                TL_ACC(w_tempty, t0);
                asm volatile("tcgen05.fence::after_thread_sync;");
                for (int kb = 0; kb < nk; ++kb, ++it) {
                    const uint32_t sa = it % NA, s = it % NB, ph = (it / NB) & 1;
                    TL_T0(t1);
                    if (PAIR) mbar_wait_cluster(&bfull[s], ph); else mbar_wait(&bfull[s], ph);   // TMA: the split B tiles
                    TL_ACC(w_full, t1);
                    TL_T0(t2);
                    if (PAIR) mbar_wait_cluster(&conv[s], ph); else mbar_wait(&conv[s], ph);    // converter warps: A lo (implies raw A landed)
                    TL_ACC(w_conv, t2);
                    asm volatile("tcgen05.fence::after_thread_sync;");
                    const uint32_t a_hi = smem_u32(smem + sa * S::A_BYTES), a_lo = smem_u32(smem + S::B_RING + s * S::BSTAGE_BYTES),
                                   b_hi = a_lo + S::A_BYTES, b_lo = b_hi + S::B_BYTES;
                    const uint32_t tacc = tmem_base + a * NACC * BN;
                    const bool first = kb == 0;  // the first k-step of a tile overwrites the accumulator
#pragma unroll
                    for (int k = 0; k < BK / 8; ++k) {
                        const uint32_t off = k * 32;  // 8 tf32 = 32 bytes inside the 128-byte swizzle row
                        if (NACC == 2) {
                            umma_tf32(tacc, desc_kmajor(a_hi + off), desc_kmajor(b_hi + off), idesc_tf32_m(BM, 2 * BN), (first && k == 0) ? 0u : 1u);
                            umma_tf32(tacc + BN, desc_kmajor(a_lo + off), desc_kmajor(b_hi + off), idesc, 1u);
                        } else if (PAIR) {
                            umma_tf32_pair(tacc, desc_kmajor(a_lo + off), desc_kmajor(b_hi + off), idesc, (first && k == 0) ? 0u : 1u);
                            umma_tf32_pair(tacc, desc_kmajor(a_hi + off), desc_kmajor(b_lo + off), idesc, 1u);
                            umma_tf32_pair(tacc, desc_kmajor(a_hi + off), desc_kmajor(b_hi + off), idesc, 1u);
                        } else {
                            umma_tf32(tacc, desc_kmajor(a_lo + off), desc_kmajor(b_hi + off), idesc, (first && k == 0) ? 0u : 1u);
                            umma_tf32(tacc, desc_kmajor(a_hi + off), desc_kmajor(b_lo + off), idesc, 1u);
                            umma_tf32(tacc, desc_kmajor(a_hi + off), desc_kmajor(b_hi + off), idesc, 1u);
                        }
                    }
                    // free both ring stages (in both CTAs when PAIR) once these MMAs have read them
                    if (PAIR) { umma_commit_pair(&aempty[sa]); umma_commit_pair(&bempty[s]); } else { umma_commit(&aempty[sa]); umma_commit(&bempty[s]); }
                }
                if (PAIR) umma_commit_pair(&tfull[a]); else umma_commit(&tfull[a]);      // accumulator complete
            }
            TL_SET(0, 3, w_tempty); TL_SET(0, 4, w_full); TL_SET(0, 5, w_conv);
        }
    } else if (warp >= CONV_T0 / 32) {
        // ===== converter warps: A lo = tf32(x - trunc13(x)), same swizzled offsets as the raw tile (elementwise) =====
        const int t = threadIdx.x - CONV_T0;
        constexpr int VEC = S::A_BYTES / 16 / (32 * CONV_WARPS);   // float4 per thread per k-block
        TL_DECL(c_wb); TL_DECL(c_wa); TL_DECL(c_work); TL_DECL(c_sig);
        uint32_t it = 0;
        for (int tile = worker; tile < tiles; tile += workers) {
            for (int kb = 0; kb < nk; ++kb, ++it) {
                const uint32_t sa = it % NA, pha = (it / NA) & 1, s = it % NB, ph = (it / NB) & 1;
                TL_T0(q0);
                mbar_wait(&bempty[s], ph ^ 1);   // the lo slot is free
                TL_ACC(c_wb, q0);
                TL_T0(q1);
                mbar_wait(&afull[sa], pha);      // the raw tile has landed
                TL_ACC(c_wa, q1);
                TL_T0(q2);
                const uint32_t raw = smem_u32(smem + sa * S::A_BYTES) + 16 * t, lo = smem_u32(smem + S::B_RING + s * S::BSTAGE_BYTES) + 16 * t;
                float4 x[VEC];
#pragma unroll
                for (int i = 0; i < VEC; ++i) x[i] = lds_v4(raw + i * 512 * CONV_WARPS);
                if (NACC == 2) {
                    // accurate variant: hi rounded to nearest IN PLACE (|lo| <= 2^-12 |x| instead of the truncation's 2^-10: the lo half's
                    // own tf32 rounding and the dropped lo*lo term shrink 4x / 16x)
#pragma unroll
                    for (int i = 0; i < VEC; ++i) {
                        const float4 h = make_float4(tf32_rna_fast(x[i].x), tf32_rna_fast(x[i].y), tf32_rna_fast(x[i].z), tf32_rna_fast(x[i].w));
                        sts_v4(raw + i * 512 * CONV_WARPS, h);
                        sts_v4(lo + i * 512 * CONV_WARPS, make_float4(tf32_rna_fast(x[i].x - h.x), tf32_rna_fast(x[i].y - h.y),
                                                                      tf32_rna_fast(x[i].z - h.z), tf32_rna_fast(x[i].w - h.w)));
                    }
                } else {
#pragma unroll
                    for (int i = 0; i < VEC; ++i) sts_v4(lo + i * 512 * CONV_WARPS, lo_trunc4(x[i]));
                }
                TL_ACC(c_work, q2);
                TL_T0(q3);
                fence_async_smem();      // generic-proxy writes -> visible to the tensor core's async-proxy reads
                if (PAIR) mbar_arrive_leader(&conv[s]); else mbar_arrive(&conv[s]);   // (per-thread arrives measured faster than elect + __syncwarp)
                TL_ACC(c_sig, q3);
            }
        }
        if (t == 0) { TL_SET(0, 8, c_wb); TL_SET(0, 9, c_wa); TL_SET(0, 10, c_work); TL_SET(0, 11, c_sig); }
    } else {
        // ===== epilogue warps: TMEM lane quarter q = warp % 4, column group cg = (warp - 2) / 4 =====
        const int q = warp & 3, cg = (warp - 2) >> 2;
        constexpr int CHUNKS = BN / 32, PER = CHUNKS / (EPI_WARPS / 4) > 0 ? CHUNKS / (EPI_WARPS / 4) : 1;
        TL_DECL(w_tfull);
        uint32_t tl = 0;
        for (int tile = worker; tile < tiles; tile += workers, ++tl) {
            const int m0 = (tile / n_tiles) * TM + (int)rank * BM, n0 = (tile % n_tiles) * BN;
            const uint32_t a = (NBUF == 2) ? (tl & 1) : 0, aph = (NBUF == 2) ? ((tl >> 1) & 1) : (tl & 1);
            const int row = m0 + q * 32 + lane;
            if (EPI == EPI_DGRAD) {
                // dgrad epilogue, software-pipelined over 16-column half-chunks: the h (post-activation) loads of half-chunk i + 1 are in
                // flight while i is processed, and those of the first half-chunk are issued BEFORE waiting for the accumulator, so
                // their HBM latency hides behind the tile's main loop
                constexpr int HC = 2 * PER;
                const bool row_ok = row < g.M;
                const size_t rbase = (size_t)(row_ok ? row : 0) * g.ldo;
                const int cbase = n0 + cg * PER * 32;
                float hn[16];
                if (cbase < g.Nout) { ldg_v8(g.aux + rbase + cbase, hn); ldg_v8(g.aux + rbase + cbase + 8, hn + 8); }
                TL_T0(t0);
                mbar_wait(&tfull[a], aph);
                TL_ACC(w_tfull, t0);
                asm volatile("tcgen05.fence::after_thread_sync;");
#pragma unroll 1
                for (int i = 0; i < HC; ++i) {
                    const int col0 = cbase + i * 16;
                    if (col0 >= g.Nout) break;               // warp-uniform
                    float hc[16];
#pragma unroll
                    for (int j = 0; j < 16; ++j) hc[j] = hn[j];
                    if (i + 1 < HC && col0 + 16 < g.Nout) { ldg_v8(g.aux + rbase + col0 + 16, hn); ldg_v8(g.aux + rbase + col0 + 24, hn + 8); }
                    uint32_t r[16];
                    tmem_ld16(tmem_base + a * NACC * BN + ((uint32_t)(q * 32) << 16) + (uint32_t)(col0 - n0), r);
                    float v[16];
#pragma unroll
                    for (int j = 0; j < 16; ++j) v[j] = __uint_as_float(r[j]) * ((hc[j] > 0.0f) ? 1.0f : (hc[j] + 1.0f));
                    if (row_ok) { stg_v8(g.out + rbase + col0, v); stg_v8(g.out + rbase + col0 + 8, v + 8); }
                    if (g.colsum) {
                        if (!row_ok) {
#pragma unroll
                            for (int j = 0; j < 16; ++j) v[j] = 0.0f;
                        }
                        // column sums over this warp's 32 rows: reduce-scatter over 16 values (15 shuffles) + the other half-warp
#pragma unroll
                        for (int half = 8; half >= 1; half >>= 1) {
                            const bool upper = (lane & half) != 0;
#pragma unroll
                            for (int j = 0; j < half; ++j) {
                                const float send = upper ? v[j] : v[j + half];
                                const float keep = upper ? v[j + half] : v[j];
                                v[j] = keep + __shfl_xor_sync(0xffffffffu, send, half);
                            }
                        }
                        const float tot = v[0] + __shfl_xor_sync(0xffffffffu, v[0], 16);
                        if (lane < 16) atomicAdd(g.colsum + col0 + lane, tot);   // lane l holds column l (bit i of l picks the upper half at step 2^i)
                    }
                }
                asm volatile("tcgen05.fence::before_thread_sync;");
                __syncwarp();
                if (lane == 0) { if (PAIR) mbar_arrive_leader(&tempty[a]); else mbar_arrive(&tempty[a]); }
                continue;
            }
            TL_T0(t0);
            mbar_wait(&tfull[a], aph);
            TL_ACC(w_tfull, t0);
            asm volatile("tcgen05.fence::after_thread_sync;");
            constexpr int nacc_used = NACC;
#pragma unroll 1
            for (int cc = 0; cc < PER; ++cc) {
                const int c = cg * PER + cc;
                if (c >= CHUNKS) break;
                const int col0 = n0 + c * 32;
                uint32_t r[32];
                tmem_ld32(tmem_base + a * NACC * BN + ((uint32_t)(q * 32) << 16) + (uint32_t)(c * 32), r);
                if (NACC > 1) {
                    for (int aa = 1; aa < nacc_used; ++aa) {
                        uint32_t r2[32];
                        tmem_ld32(tmem_base + (a * NACC + aa) * BN + ((uint32_t)(q * 32) << 16) + (uint32_t)(c * 32), r2);
#pragma unroll
                        for (int j = 0; j < 32; ++j) r[j] = __float_as_uint(__uint_as_float(r[j]) + __uint_as_float(r2[j]));
                    }
                }
                if (col0 >= g.Nout) continue;           // warp-uniform
                const bool row_ok = row < g.M;          // per lane; invalid rows compute on row 0 and contribute / store nothing
                const size_t base = (size_t)(row_ok ? row : 0) * g.ldo + col0;
                float v[32];
#pragma unroll
                for (int j = 0; j < 32; j += 4) {   // (EPI_FWD; the dgrad epilogue is the pipelined loop above)
                    const float4 b4 = __ldg(reinterpret_cast<const float4*>(g.bias + col0 + j));
                    const float bb[4] = {b4.x, b4.y, b4.z, b4.w};
#pragma unroll
                    for (int t = 0; t < 4; ++t) {
                        const float x = __uint_as_float(r[j + t]) + bb[t];
                        v[j + t] = (NACC > 1) ? ((x > 0.0f) ? x : expm1f(x)) : elu_fast(x);  // the accurate (two-accumulator) variant keeps expm1f
                    }
                }
                if (row_ok) {
#pragma unroll
                    for (int j = 0; j < 32; j += 8) stg_v8(g.out + base + j, v + j);
                }
            }
            asm volatile("tcgen05.fence::before_thread_sync;");
            __syncwarp();
            if (lane == 0) { if (PAIR) mbar_arrive_leader(&tempty[a]); else mbar_arrive(&tempty[a]); }
        }
        if (threadIdx.x == 64) TL_SET(0, 6, w_tfull);
    }
    asm volatile("tcgen05.fence::before_thread_sync;");
    if (PAIR) cluster_sync_all(); else __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;");
    if (threadIdx.x == 0) TL_SET(0, 7, tl_now());
    if (warp == 1) {
        if (PAIR) asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TCOLS));
        else asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TCOLS));
    }
}

// ---- weight gradient: D[Nout, Kin] += sum over rows m of dY[m, Nout]^T X[m, Kin] ------------------------------------
struct WgradArgs {
    float* P;         // partial tiles [tile = blockIdx.z * gridDim.y + blockIdx.y][part = blockIdx.x][128][BN] fp32 (plain coalesced
                      // stores; k_wgrad_reduce sums the parts - 148 CTAs adding into one tile with atomics cost 10-55 us per launch)
    int M, Nout, Kin; // Kin = columns stored (<= padded width covered by the tensor map boxes)
    int chunk;        // rows per CTA along blockIdx.x (multiple of 32)
    int tl_slot;      // timeline slot (debug builds only)
};

// Each CTA owns ONE output tile and a contiguous range of `chunk` rows of the reduction (so a single wave of ~148 CTAs
// covers the problem) and stores its partial tile; k_wgrad_reduce (learner_kernels.cu) sums the parts of all six weight
// gradients of an epoch in one launch.  To keep the truncating TMEM accumulation chains short, k-blocks rotate over
// NACC = 512 / BN (<= 4) accumulators that the epilogue adds with round-to-nearest FADDs.
static constexpr int WG_CONV_WARPS = 8;
static constexpr int WG_CONV_T0 = 192;                       // TMA warp, MMA warp, 4 epilogue warps, then the converters
static constexpr int WG_THREADS = WG_CONV_T0 + 32 * WG_CONV_WARPS;
template <int BN, int STAGES, int WBK>   // WBK = rows (samples) per pipeline stage: 16 or 32
__global__ void __launch_bounds__(WG_THREADS, 1)
k_tc_wgrad(const __grid_constant__ CUtensorMap mY, const __grid_constant__ CUtensorMap mX, const WgradArgs g) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    // stage layout: [dY raw -> hi (rounded in place)][X raw = hi][dY lo][X lo]
    constexpr int A_BYTES = BM * WBK * 4, B_BYTES = BN * WBK * 4, RAW_BYTES = A_BYTES + B_BYTES, STAGE_BYTES = 2 * RAW_BYTES;
    constexpr int BOX_BYTES = WBK * 128;   // one TMA box: WBK rows of 32 floats
    constexpr int NACC = (512 / BN) > 4 ? 4 : (512 / BN);
    constexpr uint32_t TCOLS = NACC * BN < 32 ? 32 : NACC * BN;
    uint64_t* full = (uint64_t*)(smem + STAGES * STAGE_BYTES);
    uint64_t* empty = full + STAGES;
    uint64_t* conv = empty + STAGES;
    uint64_t* tfull = conv + STAGES;
    uint32_t* tmem_slot = (uint32_t*)(tfull + 1);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int r_begin = blockIdx.x * g.chunk, r_end = min(g.M, r_begin + g.chunk);
    const int n0 = blockIdx.y * BM, k0 = blockIdx.z * BN;
    const int nk = (r_end > r_begin) ? (r_end - r_begin + WBK - 1) / WBK : 0;
    if (warp == 0 && lane == 0) {
        for (int s = 0; s < STAGES; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); mbar_init(&conv[s], 32 * WG_CONV_WARPS); }
        mbar_init(tfull, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(TCOLS));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    asm volatile("tcgen05.fence::before_thread_sync;");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;");
    const uint32_t tmem_base = *tmem_slot;
    if (threadIdx.x == 0) TL_SET(1, 0, tl_now());
    if (nk > 0) {
        if (warp == 0) {
            if (lane == 0) {
                TL_DECL(w_empty);
                for (int kb = 0; kb < nk; ++kb) {
                    const uint32_t s = kb % STAGES, ph = (kb / STAGES) & 1;
                    TL_T0(t0);
                    mbar_wait(&empty[s], ph ^ 1);
                    TL_ACC(w_empty, t0);
                    const uint32_t st = smem_u32(smem + s * STAGE_BYTES);
                    mbar_expect_tx(&full[s], RAW_BYTES);
                    const int r = r_begin + kb * WBK;
                    // 3-D maps (32 floats, rows, 32-float column chunks): ONE bulk copy lands all chunks of a tile, chunk c at c * BOX_BYTES
                    tma_load_3d(&mY, &full[s], st, 0, r, n0 / 32);
                    tma_load_3d(&mX, &full[s], st + A_BYTES, 0, r, k0 / 32);
                }
                TL_SET(1, 1, tl_now());
                TL_SET(1, 2, w_empty);
            }
        } else if (warp == 1) {
            if (lane == 0) {
                constexpr uint32_t idesc = idesc_tf32(BN, true);
                TL_DECL(w_full); TL_DECL(w_conv);
                for (int kb = 0; kb < nk; ++kb) {
                    const uint32_t s = kb % STAGES, ph = (kb / STAGES) & 1;
                    TL_T0(t1);
                    mbar_wait(&full[s], ph);
                    TL_ACC(w_full, t1);
                    TL_T0(t2);
                    mbar_wait(&conv[s], ph);
                    TL_ACC(w_conv, t2);
                    asm volatile("tcgen05.fence::after_thread_sync;");
                    const uint32_t a_hi = smem_u32(smem + s * STAGE_BYTES), b_hi = a_hi + A_BYTES, a_lo = a_hi + RAW_BYTES, b_lo = a_lo + A_BYTES;
                    const uint32_t tacc = tmem_base + (uint32_t)(kb % NACC) * BN;
                    const bool first = kb < NACC;  // first k-block of this accumulator overwrites it
#pragma unroll
                    for (int k = 0; k < WBK / 8; ++k) {
                        const uint32_t off = k * 1024;  // 8 rows (samples) = two 4-row swizzle groups
                        umma_tf32(tacc, desc_mnmajor(a_lo + off, BOX_BYTES), desc_mnmajor(b_hi + off, BOX_BYTES), idesc, (first && k == 0) ? 0u : 1u);
                        umma_tf32(tacc, desc_mnmajor(a_hi + off, BOX_BYTES), desc_mnmajor(b_lo + off, BOX_BYTES), idesc, 1u);
                        umma_tf32(tacc, desc_mnmajor(a_hi + off, BOX_BYTES), desc_mnmajor(b_hi + off, BOX_BYTES), idesc, 1u);
                    }
                    umma_commit(&empty[s]);
                }
                umma_commit(tfull);
                TL_SET(1, 3, tl_now()); TL_SET(1, 4, w_full); TL_SET(1, 5, w_conv);
            }
        } else if (warp >= WG_CONV_T0 / 32) {
            // ===== converter warps: split both operands in shared memory (elementwise, layout-agnostic) =====
            const int t = threadIdx.x - WG_CONV_T0;
            constexpr int NT = 32 * WG_CONV_WARPS;
            constexpr int A_VEC = A_BYTES / 16 / NT, B_VEC = B_BYTES / 16 / NT;
            static_assert(A_VEC * NT * 16 == A_BYTES && B_VEC * NT * 16 == B_BYTES, "converter tiling");
            for (int kb = 0; kb < nk; ++kb) {
                const uint32_t s = kb % STAGES, ph = (kb / STAGES) & 1;
                mbar_wait(&full[s], ph);
                const uint32_t raw = smem_u32(smem + s * STAGE_BYTES) + 16 * t, lo = raw + RAW_BYTES;
#ifndef B200_NO_CONV
                float4 x[A_VEC + B_VEC];
#pragma unroll
                for (int i = 0; i < A_VEC + B_VEC; ++i) x[i] = lds_v4(raw + i * 16 * NT);
#pragma unroll
                for (int i = 0; i < A_VEC; ++i) {   // dY: hi rounded to nearest, written back in place
                    const float4 h = make_float4(tf32_rna_fast(x[i].x), tf32_rna_fast(x[i].y), tf32_rna_fast(x[i].z), tf32_rna_fast(x[i].w));
                    sts_v4(raw + i * 16 * NT, h);
                    sts_v4(lo + i * 16 * NT, make_float4(tf32_rna_fast(x[i].x - h.x), tf32_rna_fast(x[i].y - h.y), tf32_rna_fast(x[i].z - h.z),
                                                         tf32_rna_fast(x[i].w - h.w)));
                }
#pragma unroll
                for (int i = A_VEC; i < A_VEC + B_VEC; ++i) sts_v4(lo + i * 16 * NT, lo_trunc4(x[i]));   // X: the tensor core truncates the raw word
#endif
                fence_async_smem();
                mbar_arrive(&conv[s]);
            }
        } else {
            const int q = warp & 3;
            mbar_wait(tfull, 0);
            if (threadIdx.x == 64) TL_SET(1, 6, tl_now());
            asm volatile("tcgen05.fence::after_thread_sync;");
            const int nacc_used = nk < NACC ? nk : NACC;
            float* dst = g.P + (((size_t)(blockIdx.z * gridDim.y + blockIdx.y) * gridDim.x + blockIdx.x) * BM + (q * 32 + lane)) * BN;
#pragma unroll 1
            for (int c = 0; c < BN / 32; ++c) {
                uint32_t r[32];
                float acc[32];
                tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(c * 32), r);
#pragma unroll
                for (int j = 0; j < 32; ++j) acc[j] = __uint_as_float(r[j]);
                for (int a = 1; a < nacc_used; ++a) {
                    tmem_ld32(tmem_base + (uint32_t)a * BN + ((uint32_t)(q * 32) << 16) + (uint32_t)(c * 32), r);
#pragma unroll
                    for (int j = 0; j < 32; ++j) acc[j] += __uint_as_float(r[j]);
                }
#pragma unroll
                for (int j = 0; j < 32; j += 8) stg_v8(dst + c * 32 + j, acc + j);
            }
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;");
    __syncthreads();
    if (threadIdx.x == 0) TL_SET(1, 7, tl_now());
    if (warp == 1) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TCOLS));
}

// ---- host side: tensor-map cache + launchers -------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

struct MapCache {
    EncodeTiledFn enc = nullptr;
    std::map<std::tuple<const void*, int, int, int, int, int>, CUtensorMap> maps;
    const char* error = nullptr;

    // fp32 matrix [rows, cols] with leading dimension ld (floats); box = 32 floats x box_rows; kmajor -> SWIZZLE_128B,
    // otherwise the MN-major (BASE32B) swizzle
    const CUtensorMap* get(const float* base, int rows, int cols, int ld, int box_rows, bool kmajor) {
        if (!enc) {
            void* fn = nullptr;
            cudaDriverEntryPointQueryResult q;
            if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q) != cudaSuccess || !fn) {
                error = "cuTensorMapEncodeTiled is unavailable";
                return nullptr;
            }
            enc = (EncodeTiledFn)fn;
        }
        auto key = std::make_tuple((const void*)base, rows, cols, ld, box_rows, kmajor ? 1 : 0);
        auto it = maps.find(key);
        if (it != maps.end()) return &it->second;
        CUtensorMap m;
        cuuint64_t gdim[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
        cuuint64_t gstr[1] = {(cuuint64_t)ld * 4};
        cuuint32_t box[2] = {32, (cuuint32_t)box_rows};
        cuuint32_t estr[2] = {1, 1};
        const CUresult r = enc(&m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, (void*)base, gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                               kmajor ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B,
                               CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) {
            error = "cuTensorMapEncodeTiled failed";
            return nullptr;
        }
        return &(maps[key] = m);
    }
    // MN-major operand tiles of k_tc_wgrad: the fp32 matrix [rows, cols] (cols a multiple of 32, leading dimension ld) viewed as
    // (32 floats, rows, cols / 32); box = (32, box_rows, box_chunks) with the BASE32B swizzle
    // swz128 = true: plain SWIZZLE_128B (16-bit MN-major operands: the h2 words of h2.cuh viewed as fp16 pairs)
    const CUtensorMap* get3(const float* base, int rows, int cols, int ld, int box_rows, int box_chunks, bool swz128 = false) {
        if (!enc) {
            void* fn = nullptr;
            cudaDriverEntryPointQueryResult q;
            if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q) != cudaSuccess || !fn) {
                error = "cuTensorMapEncodeTiled is unavailable";
                return nullptr;
            }
            enc = (EncodeTiledFn)fn;
        }
        auto key = std::make_tuple((const void*)base, rows, cols, ld, box_rows, (swz128 ? 2000 : 1000) + box_chunks);
        auto it = maps.find(key);
        if (it != maps.end()) return &it->second;
        CUtensorMap m;
        cuuint64_t gdim[3] = {32, (cuuint64_t)rows, (cuuint64_t)(cols / 32)};
        cuuint64_t gstr[2] = {(cuuint64_t)ld * 4, 128};
        cuuint32_t box[3] = {32, (cuuint32_t)box_rows, (cuuint32_t)box_chunks};
        cuuint32_t estr[3] = {1, 1, 1};
        const CUresult r = enc(&m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, (void*)base, gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                               swz128 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                               CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) {
            error = "cuTensorMapEncodeTiled (3-D) failed";
            return nullptr;
        }
        return &(maps[key] = m);
    }
};

template <int BN, int NA, int NB, int EPI, int NACC = 1, int PAIR = 0>
inline cudaError_t launch_rowmajor(const CUtensorMap* A, const CUtensorMap* Bh, const CUtensorMap* Bl, const RowArgs& g, int num_sms,
                                   cudaStream_t st) {
    using S = RowSmem<BN, NA, NB, PAIR>;
    static_assert(S::TOTAL <= 232448, "shared memory budget");
    static unsigned long long configured = 0;   // per-device bit mask
    {
        const cudaError_t e = ensure_dynamic_smem(k_tc_rowmajor<BN, NA, NB, EPI, NACC, PAIR>, S::TOTAL, configured);
        if (e != cudaSuccess) return e;
    }
    const int tm = PAIR ? 2 * BM : BM;
    const int tiles = ((g.M + tm - 1) / tm) * ((g.Nout + BN - 1) / BN);
    if (PAIR) {
        // clusters of two CTAs (the two SMs of a TPC): one persistent pair per two SMs
        const int pairs = tiles < num_sms / 2 ? tiles : num_sms / 2;
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3(2 * pairs);
        cfg.blockDim = dim3(ROW_THREADS);
        cfg.dynamicSmemBytes = S::TOTAL;
        cfg.stream = st;
        cudaLaunchAttribute attr[1];
        attr[0].id = cudaLaunchAttributeClusterDimension;
        attr[0].val.clusterDim.x = 2;
        attr[0].val.clusterDim.y = 1;
        attr[0].val.clusterDim.z = 1;
        cfg.attrs = attr;
        cfg.numAttrs = 1;
        return cudaLaunchKernelEx(&cfg, k_tc_rowmajor<BN, NA, NB, EPI, NACC, PAIR>, *A, *Bh, *Bl, g);
    }
    const int grid = tiles < num_sms ? tiles : num_sms;
    k_tc_rowmajor<BN, NA, NB, EPI, NACC, PAIR><<<grid, ROW_THREADS, S::TOTAL, st>>>(*A, *Bh, *Bl, g);
    return cudaPeekAtLastError();
}

template <int BN, int STAGES, int WBK>
inline cudaError_t launch_wgrad(const CUtensorMap* Y, const CUtensorMap* X, const WgradArgs& g, int kin_padded, cudaStream_t st) {
    constexpr int SMEM = STAGES * (2 * BM * WBK * 4 + 2 * BN * WBK * 4) + 256 + 1024;
    static unsigned long long configured = 0;   // per-device bit mask
    {
        const cudaError_t e = ensure_dynamic_smem(k_tc_wgrad<BN, STAGES, WBK>, SMEM, configured);
        if (e != cudaSuccess) return e;
    }
    dim3 grid((g.M + g.chunk - 1) / g.chunk, (g.Nout + BM - 1) / BM, (kin_padded + BN - 1) / BN);
    k_tc_wgrad<BN, STAGES, WBK><<<grid, WG_THREADS, SMEM, st>>>(*Y, *X, g);
    return cudaPeekAtLastError();
}

}  // namespace tc
}  // namespace b200
