// learner_kernels.cu - the PPO learner half of include/b200_t1.h: actor/critic MLPs (utils/model.py:9-36), rollout
// action sampling (utils/runner.py:109-111), GAE with time-out bootstrap (utils/utils.py:33-44, utils/runner.py:135-145),
// the full-batch epoch body with all five loss terms and their hand-derived backward (utils/runner.py:146-161), and
// clip_grad_norm_ + Adam + KL-adaptive learning rate (utils/runner.py:162-180) - all device-resident, no host sync.
//
// Dense contractions of the update go through gemm_tc.cuh (tcgen05.mma kind::tf32 with TMEM accumulators and TMA, error-
// compensated 3-term TF32 split so results stay within fp32 rounding of the reference's SGEMM; parity bar 1e-5 relative);
// the rollout policy runs on the FP32 FMA pipe (k_policy_fused); only b200_critic_value still uses the mma.sync path of
// gemm3x.cuh.  Activations live in a caller-provided workspace as plain fp32, [M, width] row-major, post-ELU
// (ELU' = h > 0 ? 1 : h + 1 needs no pre-activation copy).
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <new>

#include "common.cuh"
#include "gemm_tc.cuh"
#include "mlp_chain.cuh"
#include "h2.cuh"
#include "mlp_chain_h2.cuh"
#include "heads.cuh"
#include "rng.cuh"

using namespace b200;

// ---- flat parameter layout: state_dict tensors in the order critic.*, actor.*, logstd, each start 16-byte aligned ----
struct ParamEntry {
    const char* name;
    int offset, rows, cols;
};
static const ParamEntry kParams[] = {
    {"critic.0.weight", 0, 256, 61},      {"critic.0.bias", 15616, 256, 1},   {"critic.2.weight", 15872, 256, 256},
    {"critic.2.bias", 81408, 256, 1},     {"critic.4.weight", 81664, 128, 256}, {"critic.4.bias", 114432, 128, 1},
    {"critic.6.weight", 114560, 1, 128},  {"critic.6.bias", 114688, 1, 1},    {"actor.0.weight", 114692, 256, 47},
    {"actor.0.bias", 126724, 256, 1},     {"actor.2.weight", 126980, 128, 256}, {"actor.2.bias", 159748, 128, 1},
    {"actor.4.weight", 159876, 128, 128}, {"actor.4.bias", 176260, 128, 1},   {"actor.6.weight", 176388, 12, 128},
    {"actor.6.bias", 177924, 12, 1},      {"logstd", 177936, 1, 12},
};
enum { P_CW0 = 0, P_CB0, P_CW1, P_CB1, P_CW2, P_CB2, P_CW3, P_CB3, P_AW0, P_AB0, P_AW1, P_AB1, P_AW2, P_AB2, P_AW3, P_AB3, P_LOGSTD, P_COUNT };
#define NPARAMS_PADDED 177948

// double accumulators (caller-owned device double[32], "dstats"): the allreduce targets of a multi-GPU run
// [0..3] advantage moments (all-reduced between epoch_a and epoch_b), [4..9] loss sums + sample count (all-reduced with
// the gradient), [10] local sum of squared gradients, [16..27] d loss / d logstd.  Cleared at the start of every epoch.
enum { DS_ADV_SUM = B200_DS_ADV_SUM, DS_ADV_SUMSQ = B200_DS_ADV_SUMSQ, DS_ADV_COUNT = B200_DS_ADV_COUNT, DS_VALUE_LOSS = B200_DS_VALUE_LOSS,
       DS_ACTOR_LOSS = B200_DS_ACTOR_LOSS, DS_BOUND_LOSS = B200_DS_BOUND_LOSS, DS_ENTROPY = B200_DS_ENTROPY, DS_KL = B200_DS_KL,
       DS_SAMPLES = B200_DS_SAMPLES, DS_GRAD_SQ = B200_DS_GRAD_SQ, DS_DLOGSTD = B200_DS_DLOGSTD, DS_COUNT = B200_DS_COUNT };

// Workspace (offsets in floats).  Activations and their gradients are plain fp32 (4 B / element): the tcgen05 GEMMs split
// them into tf32 (hi, lo) pairs inside shared memory (gemm_tc.cuh).  Only the weights keep pre-split copies ("h"/"l").
#define WP_MAX_CTAS 160   // >= CTAs of one k_tc_wgrad launch (one wave: <= SM count)
struct Workspace {
    size_t Xa, Xc;                                      // packed inputs [M,64] (obs, zero padded), [M+N,64] (obs, priv, zero padded)
    size_t Xah, Xal, Xch, Xcl;                          // the same, pre-split into tf32 hi / lo (A operand of the fused chain's first layer)
    size_t C1, C2, C3;                                  // critic post-ELU activations [M+N,256],[M+N,256],[M+N,128]
    size_t A1, A2, A3;                                  // actor post-ELU activations [M,256],[M,128],[M,128]
    size_t V, MU, ADV, RET, DV, DMU;
    size_t GC3, GC2, GC1, GA3, GA2, GA1;                // dL/dz of the hidden layers: critic [M,128],[M,256],[M,256]; actor [M,128],[M,128],[M,256]
    size_t G1, G2;                                      // layer-by-layer path: gradient ping-pong [M,256] (aliases GC1, GC2)
    size_t WP[6];                                       // partial weight-gradient tiles of the six k_tc_wgrad launches of an epoch
    size_t Wc0h, Wc0l, Wc1h, Wc1l, Wc2h, Wc2l, Wa0h, Wa0l, Wa1h, Wa1l, Wa2h, Wa2l;   // split weights, K padded to 64 for layer 0
    size_t Wc1Th, Wc1Tl, Wc2Th, Wc2Tl, Wa1Th, Wa1Tl, Wa2Th, Wa2Tl;                  // transposed split weights for dgrad
    size_t LXc, LXh, LXl, L1, L2, L3, LV;                // N-row fp32 buffers of b200_critic_value
    // ---- h2 operand format (h2.cuh): two fp16 halves per 32-bit word, power-of-two scales
    size_t SC;                                          // h2::SC_COUNT floats: gradient scales of this epoch (k_finalize_loss)
    size_t ZR;                                          // zeroed by weight_prep: uint32 amax[64] (|dV|, |dmu| of k_loss), float colabs[4][256]
    size_t WP2;                                         // partial tiles of k_wgrad_h2 [CTA][128][128]
    size_t BS;                                          // S_ACT * bias of the layers whose output is handed on as h2 words: critic 1, 2, actor 1, 2 (4 x 256)
    size_t total;
};
static Workspace make_workspace(int T, int N) {
    Workspace w;
    const size_t M = (size_t)T * N, n = (size_t)N;
    size_t o = 0;
    auto take = [&](size_t cnt) { size_t r = o; o += (cnt + 255) & ~(size_t)255; return r; };  // 1 KiB aligned (TMA needs 16 B)
    const size_t Mc = M + n;  // the critic also evaluates the N post-rollout observations (last_values, utils/runner.py:133) in the same pass
    w.Xa = take(M * 64); w.Xc = take(Mc * 64);
    w.Xah = take(M * 64); w.Xal = take(M * 64); w.Xch = take(Mc * 64); w.Xcl = take(Mc * 64);
    w.C1 = take(Mc * 256); w.C2 = take(Mc * 256); w.C3 = take(Mc * 128);
    w.A1 = take(M * 256); w.A2 = take(M * 128); w.A3 = take(M * 128);
    w.V = take(Mc); w.MU = take(M * 12); w.ADV = take(M); w.RET = take(M); w.DV = take(M); w.DMU = take(M * 12);
    w.GC3 = take(M * 128); w.GC2 = take(M * 256); w.GC1 = take(M * 256);
    w.GA3 = take(M * 128); w.GA2 = take(M * 128); w.GA1 = take(M * 256);
    w.G1 = w.GC1; w.G2 = w.GC2;
    for (int j = 0; j < 6; ++j) w.WP[j] = take((size_t)WP_MAX_CTAS * 128 * 256);
    w.Wc0h = take(256 * 64); w.Wc0l = take(256 * 64); w.Wc1h = take(256 * 256); w.Wc1l = take(256 * 256);
    w.Wc2h = take(128 * 256); w.Wc2l = take(128 * 256); w.Wa0h = take(256 * 64); w.Wa0l = take(256 * 64);
    w.Wa1h = take(128 * 256); w.Wa1l = take(128 * 256); w.Wa2h = take(128 * 128); w.Wa2l = take(128 * 128);
    w.Wc1Th = take(256 * 256); w.Wc1Tl = take(256 * 256); w.Wc2Th = take(256 * 128); w.Wc2Tl = take(256 * 128);
    w.Wa1Th = take(256 * 128); w.Wa1Tl = take(256 * 128); w.Wa2Th = take(128 * 128); w.Wa2Tl = take(128 * 128);
    w.LXc = take(n * 64); w.LXh = take(n * 64); w.LXl = take(n * 64); w.L1 = take(n * 256); w.L2 = take(n * 256); w.L3 = take(n * 128);
    w.LV = take(n);
    w.SC = take(h2::SC_COUNT); w.ZR = take(64 + 4 * 256);
    w.WP2 = take((size_t)WP_MAX_CTAS * 128 * 128);
    w.BS = take(4 * 256);
    w.total = o;
    return w;
}

// ---- multi-GPU exchange over peer memory (NVLink / NVSwitch): replaces the three NCCL all-reduces of an epoch -----------------
// Every rank owns one SYMMETRIC buffer (torch.distributed._symmetric_memory: peer-mapped into all ranks of the node):
//   [0, 1 KiB)            uint32 flag[channel][source rank]   channel 0 = gradients, 1 = advantage moments; written by the peers
//   2 x slot (epoch parity) { float grads[NPARAMS_PADDED]; double loss_sums[8]; double adv_moments[4]; }
// post:   copy my contribution into my own slot, __threadfence_system, raise my flag (sequence number) in EVERY rank's buffer
// reduce: spin on my flags until every rank has posted this sequence number, then sum the peers' slots over NVLink in rank
//         order (identical result on every rank), fused with what follows (gradient norm for clip_grad_norm_)
// A slot is rewritten two posts later; a rank gets there only after a reduce that needed every peer's NEXT post, which the peer
// issues after finishing this reduce (stream order) - so two slots are enough and no second barrier is needed.
#define XCH_MAX_WORLD 16
#define XCH_FLAG_FLOATS 256
#define XCH_SLOT_FLOATS ((NPARAMS_PADDED + 16 + 8 + 63) / 64 * 64)
struct PeerX {
    unsigned long long bufs[XCH_MAX_WORLD];
    int rank, world;
};

struct B200Ppo {
    PeerX px;
    int peers;                    // 1 once b200_ppo_bind_peers succeeded
    bool last_staged = false, last_staged_h2 = false;   // b200_ppo_epoch_a: the post-rollout observations are staged (see there)
    const float *last_obs_ptr = nullptr, *last_priv_ptr = nullptr;
    unsigned int* xch_counter;    // device [4]: block counter of the post kernels, gradient / moment exchange sequence numbers, k_adam's block counter
    B200PpoConfig cfg;
    int device;
    float *params, *grads, *adam_m, *adam_v, *scalars;
    double* dstats;
    float* ws;
    unsigned long long* act_ctr;  // device counter of b200_policy_act calls (RNG step when the caller passes B200_STEP_AUTO)
    Workspace w;
    tc::MapCache* maps;           // TMA tensor maps of the (fixed) workspace buffers, built lazily on first use
    int num_sms;
    bool chain_fwd_configured = false, chain_bwd_configured = false;   // per-handle (= per-device) dynamic shared memory opt-in
    bool actor_fwd_done = false;   // epoch_a ran the actor forward together with the critic's (one fused-chain launch)
    float* P(int i) const { return params + kParams[i].offset; }
    float* G(int i) const { return grads + kParams[i].offset; }
};

// =====================================================================================================================
// small kernels
// =====================================================================================================================

__device__ __forceinline__ void split_tf32f(float x, float& hi, float& lo) {
    hi = tc::tf32_rna(x);
    lo = tc::tf32_rna(x - hi);
}

// obs [n,47] (+ priv [n,14]) -> zero-padded GEMM operands Xa [n,64] (obs), Xc [n,64] (obs, priv) as plain fp32 (operand of the
// weight-gradient GEMMs) and pre-split into tf32 hi / lo (A operand of the fused chain's first layer).  Any output may be null.
struct PackOut {
    float *Xa, *Xah, *Xal, *Xc, *Xch, *Xcl;
    int h2;   // 1: Xah / Xch receive h2 words (h2.cuh) instead of tf32 hi halves, Xal / Xcl are not written
};
__global__ void k_pack_inputs(const float* __restrict__ obs, const float* __restrict__ priv, int n, const PackOut o) {
    const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    const size_t total = (size_t)n * 64;
    if (idx >= total) return;
    const size_t r = idx >> 6;
    const int c = (int)(idx & 63);
    float v = 0.0f;
    if (c < 47) v = obs[r * 47 + c];
    else if (c < 61 && priv) v = priv[r * 14 + (c - 47)];
    const bool a = c < 47;
    if (o.h2) {
        const float wv = __uint_as_float(h2::pack_clamped(v * h2::S_X));
        if (o.Xch) o.Xch[idx] = wv;
        if (o.Xah) o.Xah[idx] = a ? wv : 0.0f;
        return;
    }
    float hi, lo;
    split_tf32f(v, hi, lo);
    if (o.Xc) o.Xc[idx] = v;
    if (o.Xch) { o.Xch[idx] = hi; o.Xcl[idx] = lo; }
    if (o.Xa) o.Xa[idx] = a ? v : 0.0f;
    if (o.Xah) { o.Xah[idx] = a ? hi : 0.0f; o.Xal[idx] = a ? lo : 0.0f; }
}

// W [rows, cols] fp32 -> split copies: K-major [rows, cols_pad] (zero padded) and, if WTh != null, transposed [cols, rows];
// all six hidden-layer matrices in ONE launch (blockIdx.y = matrix)
struct WeightPrepJob {
    const float* W;
    float *Wh, *Wl, *WTh, *WTl;
    int rows, cols, cols_pad;
    float* colabs;    // nullable: colabs[c] += |W[r, c]| (bound of the input-gradient magnitudes, k_finalize_loss)
};
struct WeightPrepJobs {
    WeightPrepJob j[6];
    const float* bias[4];   // h2: blockIdx.y == 6 writes bias_scaled[i][c] = S_ACT * bias[i][c]
    float* bias_scaled;     // [4][256]
    int bias_n[4];
    int h2;   // 1: Wh / WTh receive the h2 words of S_W * w (B'), Wl / WTl the same words with their halves swapped (B'')
};
__global__ void k_weight_prep(const WeightPrepJobs jobs) {
    if (blockIdx.y == 6) {
        const int idx = blockIdx.x * blockDim.x + threadIdx.x, i = idx >> 8, c = idx & 255;
        if (i < 4 && c < jobs.bias_n[i]) jobs.bias_scaled[i * 256 + c] = h2::S_ACT * jobs.bias[i][c];
        return;
    }
    const WeightPrepJob& q = jobs.j[blockIdx.y];
    const float* __restrict__ W = q.W;
    float* __restrict__ Wh = q.Wh;
    float* __restrict__ Wl = q.Wl;
    float* __restrict__ WTh = q.WTh;
    float* __restrict__ WTl = q.WTl;
    const int rows = q.rows, cols = q.cols, cols_pad = q.cols_pad;
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= rows * cols_pad) return;
    const int r = idx / cols_pad, c = idx - r * cols_pad;
    float hi, lo;
    if (jobs.h2) {
        const uint32_t wa = h2::pack_clamped((c < cols ? W[(size_t)r * cols + c] : 0.0f) * h2::S_W);
        hi = __uint_as_float(wa);
        lo = __uint_as_float(h2::swap_halves(wa));
    } else {
        split_tf32f(c < cols ? W[(size_t)r * cols + c] : 0.0f, hi, lo);
    }
    Wh[idx] = hi;
    Wl[idx] = lo;
    if (q.colabs && c < cols) atomicAdd(q.colabs + c, fabsf(W[(size_t)r * cols + c]));
    if (WTh && c < cols) {
        WTh[(size_t)c * rows + r] = hi;
        WTl[(size_t)c * rows + r] = lo;
    }
}

// critic head: V[m] = b + sum_k H[m,k] w[k].  Four lanes per row, eight rows per warp: every load instruction reads eight 64-byte row
// segments (whole sectors), each lane forms the partial dot product of its 32 elements and two shuffles finish it - 6 instead of
// ~20 warp instructions per row of the one-warp-per-row version.
__global__ void __launch_bounds__(256) k_value_head(const float* __restrict__ H, const float* __restrict__ w, const float* __restrict__ b, int n,
                                                    float* __restrict__ V) {
    const int gt = blockIdx.x * blockDim.x + threadIdx.x, row = gt >> 2, sub = gt & 3;
    const bool ok = row < n;
    const float4* hp = reinterpret_cast<const float4*>(H + (size_t)(ok ? row : 0) * 128) + sub;
    const float4* wp = reinterpret_cast<const float4*>(w) + sub;
    float4 h[8];
#pragma unroll
    for (int c = 0; c < 8; ++c) h[c] = ok ? __ldg(hp + 4 * c) : make_float4(0.f, 0.f, 0.f, 0.f);
    float s = 0.0f;
#pragma unroll
    for (int c = 0; c < 8; ++c) {
        const float4 ww = __ldg(wp + 4 * c);
        s = fmaf(h[c].x, ww.x, s); s = fmaf(h[c].y, ww.y, s); s = fmaf(h[c].z, ww.z, s); s = fmaf(h[c].w, ww.w, s);
    }
    s += __shfl_xor_sync(0xffffffffu, s, 1);
    s += __shfl_xor_sync(0xffffffffu, s, 2);
    if (ok && sub == 0) V[row] = s + b[0];
}

// backward of the critic head: dH[m,k] = dV[m] w[k] ELU'(H[m,k]); dw[k] += sum_m dV[m] H[m,k]; db += sum_m dV[m];
// db_prev[k] += sum_m dH[m,k] (bias gradient of the layer below).  One warp per row, one float4 (4 hidden units) per lane:
// every access is a coalesced 512-byte row; 256 rows per block, fp32 partial sums over 32 rows per thread, combined across
// the block's warps in double.
#define HB_ROWS 256
#define HB_THREADS 256
__global__ void __launch_bounds__(HB_THREADS) k_value_head_bwd(const float* __restrict__ H, const float* __restrict__ w,
                                                               const float* __restrict__ dV, int n, float* __restrict__ dH,
                                                               float* __restrict__ dw, float* __restrict__ db, float* __restrict__ db_prev,
                                                               const float* __restrict__ sg) {
    const float gs = sg ? sg[0] : 0.0f;   // != 0: dH is written as h2 words of gs * dH (h2.cuh)
    __shared__ float4 red[2][HB_THREADS / 32][32];
    __shared__ float redb[HB_THREADS / 32];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const float4 w4 = reinterpret_cast<const float4*>(w)[lane];
    const int r0 = blockIdx.x * HB_ROWS + warp * (HB_ROWS / 8), r1 = min(n, r0 + HB_ROWS / 8);
    float4 aw = make_float4(0.f, 0.f, 0.f, 0.f), ap = aw;
    float ab = 0.0f;
#pragma unroll 8
    for (int r = r0; r < r1; ++r) {
        const float g = __ldg(dV + r);
        const float4 h = reinterpret_cast<const float4*>(H + (size_t)r * 128)[lane];
        float4 d;
        d.x = g * w4.x * ((h.x > 0.0f) ? 1.0f : (h.x + 1.0f));
        d.y = g * w4.y * ((h.y > 0.0f) ? 1.0f : (h.y + 1.0f));
        d.z = g * w4.z * ((h.z > 0.0f) ? 1.0f : (h.z + 1.0f));
        d.w = g * w4.w * ((h.w > 0.0f) ? 1.0f : (h.w + 1.0f));
        if (gs != 0.0f) {
            uint4 wd;
            wd.x = h2::pack(d.x * gs); wd.y = h2::pack(d.y * gs); wd.z = h2::pack(d.z * gs); wd.w = h2::pack(d.w * gs);
            reinterpret_cast<uint4*>(dH + (size_t)r * 128)[lane] = wd;
        } else {
            reinterpret_cast<float4*>(dH + (size_t)r * 128)[lane] = d;
        }
        aw.x = fmaf(g, h.x, aw.x); aw.y = fmaf(g, h.y, aw.y); aw.z = fmaf(g, h.z, aw.z); aw.w = fmaf(g, h.w, aw.w);
        ap.x += d.x; ap.y += d.y; ap.z += d.z; ap.w += d.w;
        ab += g;
    }
    red[0][warp][lane] = aw;
    red[1][warp][lane] = ap;
    if (lane == 0) redb[warp] = ab;
    __syncthreads();
    if (threadIdx.x < 128) {
        const int k = threadIdx.x;
        double sw = 0.0, sp = 0.0;
#pragma unroll
        for (int v = 0; v < HB_THREADS / 32; ++v) {
            sw += (double)reinterpret_cast<const float*>(&red[0][v][0])[k];
            sp += (double)reinterpret_cast<const float*>(&red[1][v][0])[k];
        }
        atomicAdd(dw + k, (float)sw);
        atomicAdd(db_prev + k, (float)sp);
        if (k == 0) {
            double sb = 0.0;
            for (int v = 0; v < HB_THREADS / 32; ++v) sb += (double)redb[v];
            atomicAdd(db, (float)sb);
        }
    }
}

// actor head (utils/model.py:25): MU[m, 0..11] = b + H[m, :] W[12,128]^T, one warp per row, plain fp32 FMAs (the 12-wide
// layer is 2 % of the MLP's FLOPs: a tensor-core tile would be 90 % padding).  Each lane holds 4 hidden units and forms its 12
// partial dot products; a butterfly reduce-scatter over 16 values (15 + 1 shuffles instead of 12 x 5) leaves output j on
// lane bitrev4(j) of both half-warps.
__device__ __forceinline__ float warp_reduce_scatter16(float (&p)[16], int lane) {
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
    return p[0] + __shfl_xor_sync(0xffffffffu, p[0], 16);   // lane l (mod 16) now holds output index with bits (l&8 ? 8:0)|(l&4 ? 4:0)|...
}
// (the forward kernel below keeps W in shared memory and gives each row to FOUR lanes - 32 of its 128 hidden units each - so a row
// costs 2 x 12 shuffles instead of the 32 of a 16-value butterfly per warp-row, and every LDS.128 of W feeds two rows)
__global__ void __launch_bounds__(256) k_actor_head(const float* __restrict__ Hf, const float* __restrict__ W,
                                                    const float* __restrict__ b, int n, float* __restrict__ MU) {
    __shared__ float4 sw[12][32];
    for (int i = threadIdx.x; i < 12 * 32; i += 256) sw[i >> 5][i & 31] = reinterpret_cast<const float4*>(W)[i];
    __syncthreads();
    const int lane = threadIdx.x & 31, sub = lane & 3, rl = lane >> 2;
    const int warps = (gridDim.x * blockDim.x) >> 5;
    float bj[3];
#pragma unroll
    for (int t = 0; t < 3; ++t) bj[t] = b[3 * sub + t];
    // a warp takes 16 rows per iteration: rows base + rl and base + 8 + rl
    for (int base = (int)((blockIdx.x * blockDim.x + threadIdx.x) >> 5) * 16; base < n; base += 16 * warps) {
        const int r0 = base + rl, r1 = base + 8 + rl;
        const bool ok0 = r0 < n, ok1 = r1 < n;
        const float4* p0 = reinterpret_cast<const float4*>(Hf + (size_t)(ok0 ? r0 : 0) * 128) + sub;
        const float4* p1 = reinterpret_cast<const float4*>(Hf + (size_t)(ok1 ? r1 : 0) * 128) + sub;
        float4 h0[8], h1[8];
#pragma unroll
        for (int c = 0; c < 8; ++c) { h0[c] = __ldg(p0 + 4 * c); h1[c] = __ldg(p1 + 4 * c); }
        float a0[12], a1[12];
#pragma unroll
        for (int j = 0; j < 12; ++j) { a0[j] = 0.0f; a1[j] = 0.0f; }
#pragma unroll
        for (int c = 0; c < 8; ++c) {
#pragma unroll
            for (int j = 0; j < 12; ++j) {
                const float4 w4 = sw[j][4 * c + sub];
                a0[j] = fmaf(h0[c].x, w4.x, a0[j]); a0[j] = fmaf(h0[c].y, w4.y, a0[j]); a0[j] = fmaf(h0[c].z, w4.z, a0[j]); a0[j] = fmaf(h0[c].w, w4.w, a0[j]);
                a1[j] = fmaf(h1[c].x, w4.x, a1[j]); a1[j] = fmaf(h1[c].y, w4.y, a1[j]); a1[j] = fmaf(h1[c].z, w4.z, a1[j]); a1[j] = fmaf(h1[c].w, w4.w, a1[j]);
            }
        }
#pragma unroll
        for (int j = 0; j < 12; ++j) {
            a0[j] += __shfl_xor_sync(0xffffffffu, a0[j], 1); a0[j] += __shfl_xor_sync(0xffffffffu, a0[j], 2);
            a1[j] += __shfl_xor_sync(0xffffffffu, a1[j], 1); a1[j] += __shfl_xor_sync(0xffffffffu, a1[j], 2);
        }
        // all four lanes of a row hold the 12 sums: lane `sub` stores outputs 3 sub .. 3 sub + 2
#pragma unroll
        for (int t = 0; t < 3; ++t) {
            float v0 = 0.0f, v1 = 0.0f;
#pragma unroll
            for (int j = 0; j < 12; ++j) { if (j == 3 * sub + t) { v0 = a0[j]; v1 = a1[j]; } }
            if (ok0) MU[(size_t)r0 * 12 + 3 * sub + t] = v0 + bj[t];
            if (ok1) MU[(size_t)r1 * 12 + 3 * sub + t] = v1 + bj[t];
        }
    }
}

// backward of the actor head, fused: dH[m,k] = (sum_j dMU[m,j] W[j,k]) ELU'(H[m,k]);
// dW[j,k] += sum_m dMU[m,j] H[m,k]; db[j] += sum_m dMU[m,j]; db_prev[k] += sum_m dH[m,k].  One warp per row, one float4 of hidden
// units per lane (coalesced 512-byte rows), W and the 12 x 4 dW partials in registers; block-level combine in shared memory.
__global__ void __launch_bounds__(HB_THREADS, 2) k_actor_head_bwd(const float* __restrict__ Hf, const float* __restrict__ W,
                                                               const float* __restrict__ dMU, int n, float* __restrict__ dH,
                                                               float* __restrict__ dW, float* __restrict__ db, float* __restrict__ db_prev,
                                                               const float* __restrict__ sg) {
    const float gs = sg ? sg[0] : 0.0f;   // != 0: dH is written as h2 words of gs * dH (h2.cuh)
    __shared__ float4 red[HB_THREADS / 32][7][32];    // per-warp partials of 7 of the 13 outputs rows (12 dW rows + the dH column sums) at a time
    __shared__ float redb[HB_THREADS / 32][12];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    __shared__ float4 sw[12][32];   // W stays in shared memory (48 registers less: two blocks per SM keep more rows in flight)
    for (int i = threadIdx.x; i < 12 * 32; i += HB_THREADS) sw[i >> 5][i & 31] = reinterpret_cast<const float4*>(W)[i];
    __syncthreads();
    float4 acc[12];
#pragma unroll
    for (int j = 0; j < 12; ++j) acc[j] = make_float4(0.f, 0.f, 0.f, 0.f);
    float4 ap = make_float4(0.f, 0.f, 0.f, 0.f);
    float ab = 0.0f;   // lane j < 12 sums dMU[:, j]
    const int r0 = blockIdx.x * HB_ROWS + warp * (HB_ROWS / 8), r1 = min(n, r0 + HB_ROWS / 8);
#pragma unroll 2
    for (int r = r0; r < r1; ++r) {
        const float4 h = reinterpret_cast<const float4*>(Hf + (size_t)r * 128)[lane];
        const float4* dp = reinterpret_cast<const float4*>(dMU + (size_t)r * 12);
        const float4 d0 = __ldg(dp), d1 = __ldg(dp + 1), d2 = __ldg(dp + 2);
        const float d[12] = {d0.x, d0.y, d0.z, d0.w, d1.x, d1.y, d1.z, d1.w, d2.x, d2.y, d2.z, d2.w};
        float4 g = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
        for (int j = 0; j < 12; ++j) {
            const float4 wj = sw[j][lane];
            g.x = fmaf(d[j], wj.x, g.x); g.y = fmaf(d[j], wj.y, g.y); g.z = fmaf(d[j], wj.z, g.z); g.w = fmaf(d[j], wj.w, g.w);
            acc[j].x = fmaf(d[j], h.x, acc[j].x); acc[j].y = fmaf(d[j], h.y, acc[j].y);
            acc[j].z = fmaf(d[j], h.z, acc[j].z); acc[j].w = fmaf(d[j], h.w, acc[j].w);
        }
        float4 dh;
        dh.x = g.x * ((h.x > 0.0f) ? 1.0f : (h.x + 1.0f));
        dh.y = g.y * ((h.y > 0.0f) ? 1.0f : (h.y + 1.0f));
        dh.z = g.z * ((h.z > 0.0f) ? 1.0f : (h.z + 1.0f));
        dh.w = g.w * ((h.w > 0.0f) ? 1.0f : (h.w + 1.0f));
        if (gs != 0.0f) {
            uint4 wd;
            wd.x = h2::pack(dh.x * gs); wd.y = h2::pack(dh.y * gs); wd.z = h2::pack(dh.z * gs); wd.w = h2::pack(dh.w * gs);
            reinterpret_cast<uint4*>(dH + (size_t)r * 128)[lane] = wd;
        } else {
            reinterpret_cast<float4*>(dH + (size_t)r * 128)[lane] = dh;
        }
        ap.x += dh.x; ap.y += dh.y; ap.z += dh.z; ap.w += dh.w;
        float mine = 0.0f;
#pragma unroll
        for (int j = 0; j < 12; ++j) mine = (lane == j) ? d[j] : mine;
        ab += mine;
    }
    if (lane < 12) redb[warp][lane] = ab;
    // 13 x 128 outputs in two passes (shared-memory budget), 256 threads: thread t sums the 8 warp partials of outputs t, t + 256, ...
#pragma unroll
    for (int pass = 0; pass < 2; ++pass) {
        const int j0 = pass * 7, cnt = pass ? 6 : 7;
        if (pass) __syncthreads();
#pragma unroll
        for (int j = 0; j < 7; ++j)
            if (j < cnt) red[warp][j][lane] = (j0 + j < 12) ? acc[(j0 + j) < 12 ? (j0 + j) : 0] : ap;
        __syncthreads();
        for (int o = threadIdx.x; o < cnt * 128; o += HB_THREADS) {
            const int j = o >> 7, k = o & 127;
            double sacc = 0.0;
#pragma unroll
            for (int v = 0; v < HB_THREADS / 32; ++v) sacc += (double)reinterpret_cast<const float*>(&red[v][j][0])[k];
            if (j0 + j < 12) atomicAdd(dW + (j0 + j) * 128 + k, (float)sacc);
            else atomicAdd(db_prev + k, (float)sacc);
        }
    }
    if (threadIdx.x < 12) {
        double sb = 0.0;
        for (int v = 0; v < HB_THREADS / 32; ++v) sb += (double)redb[v][threadIdx.x];
        atomicAdd(db + threadIdx.x, (float)sb);
    }
}

__global__ void k_bump(unsigned long long* ctr) { *ctr += 1ull; }

// ---- K4: fused rollout policy (utils/runner.py:109-111; utils/model.py:18-32) -----------------------------------------
// One launch per env step: act = actor(obs) + exp(logstd) * eps for PF_ROWS observations per CTA.  All four layers run
// inside the CTA with the activations resident in shared memory (47 -> 256 -> 128 -> 128 -> 12); the weights stream
// from L2 in 16-deep k-tiles, transposed on the fly to [k][N].
// Hidden layers on the FP32 FMA pipe: at 4096 rows the layers are far too small for tcgen05 tiles, and the legacy mma.sync
// TF32 path (the first version: 3xTF32, 3072 HMMA per 16 rows) measured 47 us at 4096 envs and 86 us at 8192 - bound by the
// legacy MMA issue rate, ~34 TFLOP/s executed - where plain FFMA needs a third of the instructions' time and no operand split
// (and is exactly the reference's fp32 arithmetic class).  Thread tile: 4 rows x 8 (N = 256) or 4 (N = 128) columns; a warp
// shares its 4 rows (broadcast LDS.128 of A) and reads W as conflict-free LDS.128.  Head + sampling: one warp per row,
// fp32 FMAs + in-kernel Philox.
#define PF_ROWS 32      // (16 rows per CTA, 256 CTAs, measured 5 % slower: less reuse of each weight tile)
#define PF_RPW (PF_ROWS / 8)   // rows per warp
#define PF_KT 16      // k-tile depth
#define PF_THREADS 256
struct PolicyFusedArgs {
    const float* obs;      // [n, 47]
    const float *W0, *b0, *W1, *b1, *W2, *b2, *W3, *b3, *logstd;
    const float* eps_in;   // nullable [n, 12]
    const unsigned long long* ctr;  // nullable: device-side RNG step
    float* actions;        // [n, 12]
    float* mu_out;         // nullable [n, 12]
    uint64_t seed, step;
    int n, env_base, deterministic;
};

template <int K, int KP, int N, int LDA, int LDO>
__device__ __forceinline__ void pf_layer(const float* __restrict__ W, const float* __restrict__ bias, const float* As, float* Os,
                                         float* Ws /* [2][PF_KT][N] */) {
    // As: [32][LDA] fp32 (K valid columns, zero padded to KP); Os: [32][LDO]; W: [N][K] row-major in global memory
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    constexpr int G = N / 128;       // column groups of 4 per thread: columns g * 128 + 4 * lane + (0..3)
    constexpr int KT = KP / PF_KT;   // k-tiles
    static_assert(PF_ROWS == PF_RPW * (PF_THREADS / 32) && (N == 128 || N == 256) && KP % PF_KT == 0, "thread tiling");
    float acc[PF_RPW][4 * G];
#pragma unroll
    for (int r = 0; r < PF_RPW; ++r)
#pragma unroll
        for (int c = 0; c < 4 * G; ++c) acc[r][c] = 0.0f;
    auto load_tile = [&](int kt, float* dst) {
        // W[n][kt*16 + k8*8 .. +7] -> dst[k][n]: consecutive lanes take consecutive n (conflict-free stores); every lane reads one
        // full 32-byte sector of its row when rows are 16-byte aligned
        for (int idx = tid; idx < N * (PF_KT / 8); idx += PF_THREADS) {
            const int nrow = idx % N, k8 = (idx / N) * 8;
            const int kcol = kt * PF_KT + k8;
            float w[8];
            if (K % 4 == 0) {
                const float4 lo = __ldg(reinterpret_cast<const float4*>(W + (size_t)nrow * K + kcol));
                const float4 hi = __ldg(reinterpret_cast<const float4*>(W + (size_t)nrow * K + kcol + 4));
                w[0] = lo.x; w[1] = lo.y; w[2] = lo.z; w[3] = lo.w; w[4] = hi.x; w[5] = hi.y; w[6] = hi.z; w[7] = hi.w;
            } else {
#pragma unroll
                for (int i = 0; i < 8; ++i) w[i] = (kcol + i < K) ? __ldg(W + (size_t)nrow * K + kcol + i) : 0.0f;
            }
#pragma unroll
            for (int i = 0; i < 8; ++i) dst[(k8 + i) * N + nrow] = w[i];
        }
    };
    load_tile(0, Ws);
    __syncthreads();
    for (int kt = 0; kt < KT; ++kt) {
        const float* Wt = Ws + (kt & 1) * PF_KT * N;
        if (kt + 1 < KT) load_tile(kt + 1, Ws + ((kt + 1) & 1) * PF_KT * N);
#pragma unroll
        for (int kk = 0; kk < PF_KT; kk += 4) {
            float4 a4[PF_RPW];
#pragma unroll
            for (int r = 0; r < PF_RPW; ++r) a4[r] = *reinterpret_cast<const float4*>(As + (PF_RPW * warp + r) * LDA + kt * PF_KT + kk);
#pragma unroll
            for (int i = 0; i < 4; ++i) {
#pragma unroll
                for (int g = 0; g < G; ++g) {
                    const float4 w4 = *reinterpret_cast<const float4*>(Wt + (kk + i) * N + g * 128 + 4 * lane);
#pragma unroll
                    for (int r = 0; r < PF_RPW; ++r) {
                        const float av = (i == 0) ? a4[r].x : (i == 1) ? a4[r].y : (i == 2) ? a4[r].z : a4[r].w;
                        acc[r][4 * g + 0] = fmaf(av, w4.x, acc[r][4 * g + 0]);
                        acc[r][4 * g + 1] = fmaf(av, w4.y, acc[r][4 * g + 1]);
                        acc[r][4 * g + 2] = fmaf(av, w4.z, acc[r][4 * g + 2]);
                        acc[r][4 * g + 3] = fmaf(av, w4.w, acc[r][4 * g + 3]);
                    }
                }
            }
        }
        __syncthreads();
    }
    // bias + ELU -> next activation buffer
#pragma unroll
    for (int g = 0; g < G; ++g) {
        const float4 b4 = __ldg(reinterpret_cast<const float4*>(bias + g * 128 + 4 * lane));
#pragma unroll
        for (int r = 0; r < PF_RPW; ++r) {
            float4 v;
            v.x = acc[r][4 * g + 0] + b4.x; v.y = acc[r][4 * g + 1] + b4.y; v.z = acc[r][4 * g + 2] + b4.z; v.w = acc[r][4 * g + 3] + b4.w;
            v.x = (v.x > 0.f) ? v.x : expm1f(v.x);
            v.y = (v.y > 0.f) ? v.y : expm1f(v.y);
            v.z = (v.z > 0.f) ? v.z : expm1f(v.z);
            v.w = (v.w > 0.f) ? v.w : expm1f(v.w);
            *reinterpret_cast<float4*>(Os + (PF_RPW * warp + r) * LDO + g * 128 + 4 * lane) = v;
        }
    }
    __syncthreads();
}

#define PF_LD0 68    // 64 + 4
#define PF_LD1 260   // 256 + 4
#define PF_LD2 132   // 128 + 4
// two activation buffers, reused: P holds H1 then H3, Q holds X0 then H2
#define PF_SMEM_FLOATS (PF_ROWS * (PF_LD1 + PF_LD2) + 2 * PF_KT * 256 + 12 * 128)
__global__ void __launch_bounds__(PF_THREADS) k_policy_fused(const PolicyFusedArgs a) {
    extern __shared__ __align__(16) float pf_smem[];
    float* X0 = pf_smem;
    float* H2 = X0;                        // Q: X0 (ld 68) is dead once layer 1 has run
    float* H1 = X0 + PF_ROWS * PF_LD2;     // P
    float* H3 = H1;                        // P: H1 is dead once layer 2 has run
    float* Ws = H1 + PF_ROWS * PF_LD1;
    float* W3s = Ws + 2 * PF_KT * 256;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int row0 = blockIdx.x * PF_ROWS;
    for (int idx = tid; idx < PF_ROWS * 64; idx += PF_THREADS) {
        const int r = idx >> 6, c = idx & 63;
        const int row = row0 + r;
        X0[r * PF_LD0 + c] = (c < 47 && row < a.n) ? a.obs[(size_t)row * 47 + c] : 0.0f;
    }
    for (int idx = tid; idx < 12 * 128; idx += PF_THREADS) W3s[idx] = a.W3[idx];
    __syncthreads();
    pf_layer<47, 64, 256, PF_LD0, PF_LD1>(a.W0, a.b0, X0, H1, Ws);
    pf_layer<256, 256, 128, PF_LD1, PF_LD2>(a.W1, a.b1, H1, H2, Ws);
    pf_layer<128, 128, 128, PF_LD2, PF_LD2>(a.W2, a.b2, H2, H3, Ws);
    // head + sampling: warp w handles rows w, w + 8, ...
    const uint64_t step = a.ctr ? (uint64_t)(*a.ctr) : a.step;
    for (int r = warp; r < PF_ROWS; r += PF_THREADS / 32) {
        const int row = row0 + r;
        if (row >= a.n) continue;   // warp-uniform
        const float4 h = *reinterpret_cast<const float4*>(H3 + r * PF_LD2 + 4 * lane);
        float out = 0.0f;
#pragma unroll
        for (int j = 0; j < 12; ++j) {
            const float4 w = *reinterpret_cast<const float4*>(W3s + j * 128 + 4 * lane);
            float s = h.x * w.x + h.y * w.y + h.z * w.z + h.w * w.w;
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
            if (lane == j) out = s + a.b3[j];
        }
        if (lane < 12) {
            float act = out;
            if (!a.deterministic) {
                float eps;
                if (a.eps_in) eps = a.eps_in[(size_t)row * 12 + lane];
                else eps = rand4(rng_words(a.seed, (uint32_t)(a.env_base + row), step, RP_POLICY, lane >> 2)).n[lane & 3];
                act = __fadd_rn(out, __fmul_rn(expf(a.logstd[lane]), eps));
            }
            a.actions[(size_t)row * 12 + lane] = act;
            if (a.mu_out) a.mu_out[(size_t)row * 12 + lane] = out;
        }
    }
}

// Tensor-core act path (b200_policy_act for >= g_policy_tc_min_rows rows): the actor's hidden layers run on the fused forward chain
// (k_mlp_fwd_h2), this is the head + sampling stage of k_policy_fused over the chain's h3 rows - same dot-product order, same Philox
// keys (seed, global env, step, lane / 4), so both paths draw the same noise.  One warp per row, the head weights in registers.
__global__ void __launch_bounds__(256) k_act_head(const float* __restrict__ H3, const PolicyFusedArgs a) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    float4 w[12];
#pragma unroll
    for (int j = 0; j < 12; ++j) w[j] = __ldg(reinterpret_cast<const float4*>(a.W3 + j * 128) + lane);
    const float bj = lane < 12 ? a.b3[lane] : 0.0f;
    const float sg = lane < 12 ? expf(a.logstd[lane]) : 0.0f;
    const uint64_t step = a.ctr ? (uint64_t)(*a.ctr) : a.step;
    for (int row = blockIdx.x * 8 + warp; row < a.n; row += gridDim.x * 8) {
        const float4 h = __ldg(reinterpret_cast<const float4*>(H3 + (size_t)row * 128) + lane);
        float out = 0.0f;
#pragma unroll
        for (int j = 0; j < 12; ++j) {
            float s = h.x * w[j].x + h.y * w[j].y + h.z * w[j].z + h.w * w[j].w;
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
            if (lane == j) out = s + bj;
        }
        if (lane < 12) {
            float act = out;
            if (!a.deterministic) {
                float eps;
                if (a.eps_in) eps = a.eps_in[(size_t)row * 12 + lane];
                else eps = rand4(rng_words(a.seed, (uint32_t)(a.env_base + row), step, RP_POLICY, lane >> 2)).n[lane & 3];
                act = __fadd_rn(out, __fmul_rn(sg, eps));
            }
            a.actions[(size_t)row * 12 + lane] = act;
            if (a.mu_out) a.mu_out[(size_t)row * 12 + lane] = out;
        }
    }
}

#define LOG_SQRT_2PI 0.91893853320467274178f

// Normal(mu, exp(logstd)).log_prob(a).sum(-1)  (torch.distributions.Normal.log_prob, summed left to right)
__device__ __forceinline__ float normal_logp12(const float* a, const float* mu, const float* sigma, const float* log_sigma) {
    float lp = 0.0f;
#pragma unroll
    for (int j = 0; j < 12; ++j) {
        const float d = __fsub_rn(a[j], mu[j]);
        const float var = __fmul_rn(sigma[j], sigma[j]);
        const float t = __fsub_rn(__fsub_rn(__fdiv_rn(-__fmul_rn(d, d), __fmul_rn(2.0f, var)), log_sigma[j]), LOG_SQRT_2PI);
        lp = __fadd_rn(lp, t);
    }
    return lp;
}

// utils/runner.py:123-125: log-prob of the stored actions under the pre-update policy; also snapshots logstd
__global__ void k_old_logp(const float* __restrict__ mu, const float* __restrict__ actions, const float* __restrict__ logstd,
                           int M, float* __restrict__ old_mu, float* __restrict__ old_logp, float* __restrict__ scalars) {
    const int m = blockIdx.x * blockDim.x + threadIdx.x;
    if (m == 0) {
#pragma unroll
        for (int j = 0; j < 12; ++j) scalars[B200_SC_OLD_LOGSTD + j] = logstd[j];
    }
    if (m >= M) return;
    float a[12], u[12], sg[12], ls[12];
#pragma unroll
    for (int j = 0; j < 12; ++j) {
        a[j] = actions[(size_t)m * 12 + j];
        u[j] = mu[(size_t)m * 12 + j];
        sg[j] = expf(logstd[j]);
        ls[j] = logf(sg[j]);
        old_mu[(size_t)m * 12 + j] = u[j];
    }
    old_logp[m] = normal_logp12(a, u, sg, ls);
}

// utils/runner.py:135-144 + utils/utils.py:33-44: one thread per env walks t = T-1..0.
//   rewards[time_outs] = values[time_outs] (in place); done = done | time_out;
//   delta = r + gamma*nnt*V[t+1] - V[t];  A[t] = delta + gamma*lam*nnt*A[t+1];  returns = V + A
// and accumulates sum / sum of squares / count of the raw advantages (double) for the normalisation of :145.
#define GAE_CHUNK 24
__global__ void __launch_bounds__(128) k_gae(float* __restrict__ rewards, const uint8_t* __restrict__ dones,
                                             const uint8_t* __restrict__ time_outs, const float* __restrict__ values,
                                             const float* __restrict__ last_values, float gamma, float gamma_lam, int T,
                                             int N, float* __restrict__ adv, float* __restrict__ ret,
                                             double* __restrict__ stats) {
    const int e = blockIdx.x * blockDim.x + threadIdx.x;
    double s = 0.0, s2 = 0.0;
    if (e < N) {
        float next_v = last_values[e], last_adv = 0.0f;
        // the recurrence is sequential in t, its loads are not: fetch 24 time steps (the shipped horizon) at once, then run them
        for (int t1 = T; t1 > 0; t1 -= GAE_CHUNK) {
            float vv[GAE_CHUNK], rr[GAE_CHUNK];
            uint8_t dd[GAE_CHUNK], oo[GAE_CHUNK];
#pragma unroll
            for (int k = 0; k < GAE_CHUNK; ++k) {
                const int t = t1 - 1 - k;
                const size_t i = (size_t)(t >= 0 ? t : 0) * N + e;
                vv[k] = values[i]; rr[k] = rewards[i]; dd[k] = dones[i]; oo[k] = time_outs[i];
            }
            // straight-line recurrence (no early exit: the steps of a short last chunk are predicated off), so that all 4 x 24 loads
            // above are issued before the first use - with a `break` per step the compiler sank them next to their uses: 24
            // dependent memory round trips, measured 15 us for 2 MB
#pragma unroll
            for (int k = 0; k < GAE_CHUNK; ++k) {
                const int t = t1 - 1 - k;
                const bool live = t >= 0;
                const size_t i = (size_t)(live ? t : 0) * N + e;
                const float v = vv[k];
                const bool to = oo[k] != 0;
                float r = rr[k];
                if (to) r = v;
                if (live && to) rewards[i] = v;
                const float nnt = (dd[k] != 0 || to) ? 0.0f : 1.0f;
                const float delta = __fsub_rn(__fadd_rn(r, __fmul_rn(__fmul_rn(gamma, nnt), next_v)), v);
                const float a_new = __fadd_rn(delta, __fmul_rn(__fmul_rn(gamma_lam, nnt), last_adv));
                if (live) {
                    last_adv = a_new;
                    adv[i] = a_new;
                    ret[i] = __fadd_rn(v, a_new);
                    s += (double)a_new;
                    s2 += (double)a_new * (double)a_new;
                    next_v = v;
                }
            }
        }
    }
    if (stats) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            s += __shfl_xor_sync(0xffffffffu, s, o);
            s2 += __shfl_xor_sync(0xffffffffu, s2, o);
        }
        __shared__ double sh[2][4];
        const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
        if (lane == 0) { sh[0][warp] = s; sh[1][warp] = s2; }
        __syncthreads();
        if (threadIdx.x == 0) {
            double a = 0, b = 0;
            for (int w = 0; w < (int)(blockDim.x >> 5); ++w) { a += sh[0][w]; b += sh[1][w]; }
            atomicAdd(stats + DS_ADV_SUM, a);
            atomicAdd(stats + DS_ADV_SUMSQ, b);
            if (blockIdx.x == 0) atomicAdd(stats + DS_ADV_COUNT, (double)T * (double)N);
        }
    }
}

// Loss terms of utils/runner.py:146-161 and their gradients w.r.t. V, mu and logstd, one thread per sample.
//   value_loss = mean((V - returns)^2)                      dV   = 2 (V - returns) / M
//   ratio = exp(logp - old_logp); actor = mean(max(-A ratio, -A clamp(ratio, 1-e, 1+e)))   (utils/utils.py:47-52)
//   bound = mean(relu(mu - 1)^2) + mean(min(mu + 1, 0)^2) over M*12 elements
//   entropy = sum_j(0.5 + 0.5 log 2pi + logstd_j);  loss += entropy_coef * mean(entropy)
//   kl = sum_j [log(sigma/sigma_old) + (sigma_old^2 + (mu - mu_old)^2) / (2 sigma^2) - 0.5]    (no gradient)
#define LOSS_BLOCK 256
#define LOSS_NRED (6 + 12)
__global__ void __launch_bounds__(LOSS_BLOCK) k_loss(const float* __restrict__ V, const float* __restrict__ ret,
                                                     const float* __restrict__ adv, const float* __restrict__ mu,
                                                     const float* __restrict__ actions, const float* __restrict__ old_mu,
                                                     const float* __restrict__ old_logp, const float* __restrict__ logstd,
                                                     const float* __restrict__ scalars, double* __restrict__ dstats, int M,
                                                     float e_clip, float bound_coef, float* __restrict__ dV,
                                                     float* __restrict__ dMU, unsigned int* __restrict__ amax) {
    const int m = blockIdx.x * LOSS_BLOCK + threadIdx.x;
    float mx_v = 0.0f, mx_mu = 0.0f;   // largest |dV|, |dmu| of this thread (h2 gradient scale, k_finalize_loss)
    const double cnt = dstats[DS_ADV_COUNT];
    const double mean_d = dstats[DS_ADV_SUM] / cnt;
    const double var_d = fmax((dstats[DS_ADV_SUMSQ] - dstats[DS_ADV_SUM] * mean_d) / (cnt - 1.0), 0.0);
    const float a_mean = (float)mean_d, a_std = (float)sqrt(var_d);
    const float invM = 1.0f / (float)M;
    float red[LOSS_NRED];
#pragma unroll
    for (int i = 0; i < LOSS_NRED; ++i) red[i] = 0.0f;
    // per-action constants of the two distributions: the same for every sample, so 12 threads of the block evaluate them once (the
    // exp / log pairs were ~40 % of this kernel's instructions when every thread computed them for itself); same expressions, same bits
    __shared__ float c_sg[12], c_ls[12], c_sgo[12], c_klc[12];
    if (threadIdx.x < 12) {
        const float sgj = expf(logstd[threadIdx.x]), sgo = expf(scalars[B200_SC_OLD_LOGSTD + threadIdx.x]);
        c_sg[threadIdx.x] = sgj;
        c_ls[threadIdx.x] = logf(sgj);
        c_sgo[threadIdx.x] = sgo;
        c_klc[threadIdx.x] = logf(sgj / sgo);
    }
    __syncthreads();
    if (m < M) {
        float sg[12], ls[12], sg_old[12], a[12], u[12], uo[12];
        {
            // the three 48-byte rows of this sample as 3 x 3 float4 loads issued together (the old_mu reads used to sit inside the loop
            // below, one dependent memory round trip per action behind the division calls)
            const float4* a4 = reinterpret_cast<const float4*>(actions + (size_t)m * 12);
            const float4* u4 = reinterpret_cast<const float4*>(mu + (size_t)m * 12);
            const float4* o4 = reinterpret_cast<const float4*>(old_mu + (size_t)m * 12);
#pragma unroll
            for (int q = 0; q < 3; ++q) {
                const float4 x = a4[q], y = u4[q], z = o4[q];
                a[4 * q] = x.x; a[4 * q + 1] = x.y; a[4 * q + 2] = x.z; a[4 * q + 3] = x.w;
                u[4 * q] = y.x; u[4 * q + 1] = y.y; u[4 * q + 2] = y.z; u[4 * q + 3] = y.w;
                uo[4 * q] = z.x; uo[4 * q + 1] = z.y; uo[4 * q + 2] = z.z; uo[4 * q + 3] = z.w;
            }
        }
#pragma unroll
        for (int j = 0; j < 12; ++j) {
            sg[j] = c_sg[j];
            ls[j] = c_ls[j];
            sg_old[j] = c_sgo[j];
        }
        // value loss
        const float v = V[m], rt = ret[m];
        const float dv = __fsub_rn(v, rt);
        red[0] = dv * dv;
        dV[m] = 2.0f * dv * invM;
        mx_v = fabsf(2.0f * dv * invM);
        // surrogate
        const float A = __fdiv_rn(__fsub_rn(adv[m], a_mean), __fadd_rn(a_std, 1.0e-8f));
        const float lp = normal_logp12(a, u, sg, ls);
        const float ratio = expf(__fsub_rn(lp, old_logp[m]));
        const float lo = 1.0f - e_clip, hi = 1.0f + e_clip;
        const float rc = fminf(fmaxf(ratio, lo), hi);
        const float s1 = -A * ratio, s2 = -A * rc;
        red[1] = fmaxf(s1, s2);
        float dr;  // d surrogate / d ratio  (torch.max splits ties evenly; clamp passes gradient on [lo, hi])
        const bool inside = (ratio >= lo) && (ratio <= hi);
        if (inside) dr = -A;
        else if (s1 > s2) dr = -A;
        else if (s1 < s2) dr = 0.0f;
        else dr = -0.5f * A;
        const float dlp = dr * ratio * invM;
        // bound loss, entropy, kl, gradients
        float bsum = 0.0f, ent = 0.0f, kl = 0.0f, dmu[12];
        const float bscale = bound_coef * 2.0f / ((float)M * 12.0f);
#pragma unroll
        for (int j = 0; j < 12; ++j) {
            const float d = a[j] - u[j];
            const float var = sg[j] * sg[j];
            const float up = fmaxf(u[j] - 1.0f, 0.0f), dn = fminf(u[j] + 1.0f, 0.0f);
            bsum += up * up + dn * dn;
            ent += 0.5f + LOG_SQRT_2PI + ls[j];
            const float dm = u[j] - uo[j];
            kl += c_klc[j] + 0.5f * (sg_old[j] * sg_old[j] + dm * dm) / var - 0.5f;
            const float dmu_j = dlp * d / var + bscale * (up + dn);
            dmu[j] = dmu_j;
            mx_mu = fmaxf(mx_mu, fabsf(dmu_j));
            red[6 + j] = dlp * (d * d / var - 1.0f);
        }
#pragma unroll
        for (int q = 0; q < 3; ++q)
            reinterpret_cast<float4*>(dMU + (size_t)m * 12)[q] = make_float4(dmu[4 * q], dmu[4 * q + 1], dmu[4 * q + 2], dmu[4 * q + 3]);
        red[2] = bsum;
        red[3] = ent;
        red[4] = kl;
        red[5] = 1.0f;
    } else {
        // nothing: contributes zeros
    }
    // block reduction (shuffle, then shared) -> one double atomic per quantity per block
    __shared__ float sh[LOSS_BLOCK / 32][LOSS_NRED];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
#pragma unroll
    for (int i = 0; i < LOSS_NRED; ++i) {
        float x = red[i];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) x += __shfl_xor_sync(0xffffffffu, x, o);
        if (lane == 0) sh[warp][i] = x;
    }
    if (amax) {
        // non-negative floats order like their bit patterns; NaN / Inf are left out (the scale falls back to 1)
        unsigned bv = isfinite(mx_v) ? __float_as_uint(mx_v) : 0u, bm = isfinite(mx_mu) ? __float_as_uint(mx_mu) : 0u;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            bv = max(bv, __shfl_xor_sync(0xffffffffu, bv, o));
            bm = max(bm, __shfl_xor_sync(0xffffffffu, bm, o));
        }
        if (lane == 0) { if (bv) atomicMax(amax + 0, bv); if (bm) atomicMax(amax + 1, bm); }
    }
    __syncthreads();
    if (threadIdx.x < LOSS_NRED) {
        double s = 0.0;
        for (int w = 0; w < LOSS_BLOCK / 32; ++w) s += (double)sh[w][threadIdx.x];
        const int i = threadIdx.x;
        const int slot = (i < 6) ? (DS_VALUE_LOSS + i) : (DS_DLOGSTD + (i - 6));
        atomicAdd(dstats + slot, s);
    }
}


__device__ __forceinline__ unsigned ld_acquire_sys_u32(const unsigned* p) {
    unsigned v;
    asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_release_sys_u32(unsigned* p, unsigned v) {
    asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ float4 ld_peer_f4(const float4* p) {   // peer (NVLink) load, not cached in L1
    float4 v;
    asm volatile("ld.volatile.global.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p));
    return v;
}
__device__ __forceinline__ double ld_peer_f64(const double* p) {
    double v;
    asm volatile("ld.volatile.global.f64 %0, [%1];" : "=d"(v) : "l"(p));
    return v;
}
__device__ __forceinline__ float* xch_slot(const PeerX& x, int r, int parity) {
    return reinterpret_cast<float*>(x.bufs[r]) + XCH_FLAG_FLOATS + (size_t)parity * XCH_SLOT_FLOATS;
}
__device__ __forceinline__ void xch_raise(const PeerX& x, int channel, unsigned seq) {   // one thread
    for (int r = 0; r < x.world; ++r) st_release_sys_u32(reinterpret_cast<unsigned*>(x.bufs[r]) + channel * XCH_MAX_WORLD + x.rank, seq);
}
__device__ __forceinline__ void xch_wait(const PeerX& x, int channel, unsigned seq) {     // threads [0, world) of a block, then __syncthreads
    if ((int)threadIdx.x < x.world) {
        const unsigned* f = reinterpret_cast<const unsigned*>(x.bufs[x.rank]) + channel * XCH_MAX_WORLD + threadIdx.x;
        while ((int)(ld_acquire_sys_u32(f) - seq) < 0) { }
    }
    __syncthreads();
}

// gradients + loss sums -> my slot; the last block to finish raises the flags
// The sequence numbers live in device memory (counter[1] = gradient exchanges, counter[2] = moment exchanges done so far), not in kernel
// arguments: a CUDA graph of the update replays with the arguments it was captured with, and the protocol must keep counting.
__global__ void __launch_bounds__(256) k_xchg_post_grads(const PeerX x, const float* __restrict__ grads, const double* __restrict__ dstats,
                                                         unsigned* __restrict__ counter) {
    const unsigned seq = counter[1] + 1u;     // (written only by the last block, after every block has read it: see below)
    const int parity = (int)(seq & 1u);
    float* slot = xch_slot(x, x.rank, parity);
    const float4* g4 = reinterpret_cast<const float4*>(grads);
    float4* s4 = reinterpret_cast<float4*>(slot);
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < NPARAMS_PADDED / 4; i += gridDim.x * blockDim.x) s4[i] = g4[i];
    if (blockIdx.x == 0 && threadIdx.x < 6) reinterpret_cast<double*>(slot + NPARAMS_PADDED)[threadIdx.x] = dstats[DS_VALUE_LOSS + threadIdx.x];
    __threadfence_system();
    __syncthreads();
    if (threadIdx.x == 0) {
        const unsigned done = atomicAdd(counter, 1u);   // (every block read counter[1] before its arrival here)
        if (done == gridDim.x - 1) {
            counter[0] = 0u;
            counter[1] = seq;
            __threadfence_system();
            xch_raise(x, 0, seq);
        }
    }
}
// all-reduce (sum) of the gradient and the loss sums over the peers' slots, fused with the gradient norm of clip_grad_norm_
__global__ void __launch_bounds__(256) k_xchg_reduce_grads(const PeerX x, float* __restrict__ grads, double* __restrict__ dstats,
                                                           const unsigned* __restrict__ counter, float inv_world) {
    const unsigned seq = counter[1];          // the post kernel before this launch (same stream) has finished
    const int parity = (int)(seq & 1u);
    xch_wait(x, 0, seq);
    float4* g4 = reinterpret_cast<float4*>(grads);
    float s = 0.0f;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < NPARAMS_PADDED / 4; i += gridDim.x * blockDim.x) {
        float4 a = make_float4(0.f, 0.f, 0.f, 0.f);
        for (int r = 0; r < x.world; ++r) {
            const float4 v = ld_peer_f4(reinterpret_cast<const float4*>(xch_slot(x, r, parity)) + i);
            a.x += v.x; a.y += v.y; a.z += v.z; a.w += v.w;
        }
        g4[i] = a;
        const float gx = a.x * inv_world, gy = a.y * inv_world, gz = a.z * inv_world, gw = a.w * inv_world;
        s += gx * gx + gy * gy + gz * gz + gw * gw;
    }
    if (blockIdx.x == 0 && threadIdx.x < 6) {
        double t = 0.0;
        for (int r = 0; r < x.world; ++r) t += ld_peer_f64(reinterpret_cast<const double*>(xch_slot(x, r, parity) + NPARAMS_PADDED) + threadIdx.x);
        dstats[DS_VALUE_LOSS + threadIdx.x] = t;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    __shared__ float sh[8];
    if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = s;
    __syncthreads();
    if (threadIdx.x == 0) {
        double t = 0.0;
        for (int w = 0; w < 8; ++w) t += (double)sh[w];
        atomicAdd(dstats + DS_GRAD_SQ, t);
    }
}
// advantage moments (sum, sum of squares, count) of this rank -> my slot, flags raised
__global__ void k_xchg_post_stats(const PeerX x, const double* __restrict__ dstats, unsigned* __restrict__ counter) {
    const unsigned seq = counter[2] + 1u;
    const int parity = (int)(seq & 1u);
    double* d = reinterpret_cast<double*>(xch_slot(x, x.rank, parity) + NPARAMS_PADDED + 16);
    if (threadIdx.x < 4) d[threadIdx.x] = dstats[threadIdx.x];
    __threadfence_system();
    __syncthreads();
    if (threadIdx.x == 0) { counter[2] = seq; xch_raise(x, 1, seq); }   // (single block: every thread read counter[2] before the barrier)
}
__global__ void k_xchg_reduce_stats(const PeerX x, double* __restrict__ dstats, const unsigned* __restrict__ counter) {
    const unsigned seq = counter[2];
    const int parity = (int)(seq & 1u);
    xch_wait(x, 1, seq);
    if (threadIdx.x < 4) {
        double t = 0.0;
        for (int r = 0; r < x.world; ++r) t += ld_peer_f64(reinterpret_cast<const double*>(xch_slot(x, r, parity) + NPARAMS_PADDED + 16) + threadIdx.x);
        dstats[threadIdx.x] = t;
    }
}

// logstd gradient = reduced surrogate part + entropy_coef * d mean(entropy) / d logstd_j (= entropy_coef)
// + the h2 gradient scales of this epoch (h2.cuh): a power of two per net such that a RIGOROUS bound of every dL/dz magnitude of
// the net maps below 2^15 (fp16 max 65504):  |dz3| <= max|dhead| * max_k sum_j |W_head[j,k]|,  |dz2| <= |dz3|_max * max_k sum_n |W3[n,k]|
// (ELU' <= 1),  |dz1| likewise with W2.  max|dV|, max|dmu| come from k_loss, the column abs-sums from k_weight_prep.
__global__ void __launch_bounds__(256) k_finalize_loss(const double* __restrict__ dstats, float entropy_coef, float* __restrict__ g_logstd,
                                                       const float* __restrict__ Wa3, const float* __restrict__ wc3,
                                                       const unsigned int* __restrict__ amax, const float* __restrict__ cab,
                                                       float* __restrict__ sc) {
    const int t = threadIdx.x;
    if (t < 12) g_logstd[t] = (float)dstats[DS_DLOGSTD + t] + entropy_coef;
    if (!sc) return;
    float v[6] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f};   // actor head, critic head, Wc3, Wc2, Wa3, Wa2 (column abs-sums)
    if (t < 128) {
        float s = 0.0f;
        for (int j = 0; j < 12; ++j) s += fabsf(Wa3[j * 128 + t]);
        v[0] = s;
        v[1] = fabsf(wc3[t]);
        v[4] = cab[2 * 256 + t];
    }
    v[2] = cab[0 * 256 + t];
    v[3] = cab[1 * 256 + t];
    v[5] = cab[3 * 256 + t];
    __shared__ float red[8][6];
#pragma unroll
    for (int i = 0; i < 6; ++i) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v[i] = fmaxf(v[i], __shfl_xor_sync(0xffffffffu, v[i], o));
        if ((t & 31) == 0) red[t >> 5][i] = v[i];
    }
    __syncthreads();
    if (t == 0) {
        for (int i = 0; i < 6; ++i)
            for (int w = 1; w < 8; ++w) v[i] = fmaxf(v[i], red[w][i]);
        const float bound_c = __uint_as_float(amax[0]) * v[1] * fmaxf(1.0f, fmaxf(v[2], v[2] * v[3]));
        const float bound_a = __uint_as_float(amax[1]) * v[0] * fmaxf(1.0f, fmaxf(v[4], v[4] * v[5]));
        const float b[2] = {bound_c, bound_a};
        for (int i = 0; i < 2; ++i) {
            int e = 0;
            if (b[i] > 0.0f && isfinite(b[i])) e = 14 - ilogbf(b[i]);   // bound * 2^e < 2^15
            e = max(-100, min(100, e));
            sc[2 * i] = ldexpf(1.0f, e);
            sc[2 * i + 1] = ldexpf(1.0f, -e);
        }
    }
}

// clip_grad_norm_ part 1 (utils/runner.py:164): sum of squares of the (world-averaged) flat gradient
__global__ void __launch_bounds__(256) k_grad_sumsq(const float* __restrict__ grads, int n, float inv_world,
                                                    double* __restrict__ dstats) {
    float s = 0.0f;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const float g = grads[i] * inv_world;
        s += g * g;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    __shared__ float sh[8];
    if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = s;
    __syncthreads();
    if (threadIdx.x == 0) {
        double t = 0.0;
        for (int w = 0; w < 8; ++w) t += (double)sh[w];
        atomicAdd(dstats + DS_GRAD_SQ, t);
    }
}

// epoch scalars + KL-adaptive learning rate (utils/runner.py:167-184), one thread
__device__ void post_apply_body(float* __restrict__ scalars, const double* __restrict__ dstats, float desired_kl, float lr_min,
                                float lr_max, float lr_factor) {
    const double S = dstats[DS_SAMPLES];
    const float vl = (float)(dstats[DS_VALUE_LOSS] / S), al = (float)(dstats[DS_ACTOR_LOSS] / S);
    const float bl = (float)(dstats[DS_BOUND_LOSS] / (S * 12.0)), en = (float)(dstats[DS_ENTROPY] / S);
    const float kl = (float)(dstats[DS_KL] / S);
    scalars[B200_SC_VALUE_LOSS] = vl;
    scalars[B200_SC_ACTOR_LOSS] = al;
    scalars[B200_SC_BOUND_LOSS] = bl;
    scalars[B200_SC_ENTROPY] = en;
    scalars[B200_SC_KL] = kl;
    const double cnt = dstats[DS_ADV_COUNT], mean = dstats[DS_ADV_SUM] / cnt;
    scalars[B200_SC_ADV_MEAN] = (float)mean;
    scalars[B200_SC_ADV_STD] = (float)sqrt(fmax((dstats[DS_ADV_SUMSQ] - dstats[DS_ADV_SUM] * mean) / (cnt - 1.0), 0.0));
    scalars[B200_SC_GRAD_NORM] = (float)sqrt(dstats[DS_GRAD_SQ]);
    scalars[B200_SC_SUM_VALUE_LOSS] += vl;
    scalars[B200_SC_SUM_ACTOR_LOSS] += al;
    scalars[B200_SC_SUM_BOUND_LOSS] += bl;
    scalars[B200_SC_SUM_ENTROPY] += en;
    scalars[B200_SC_EPOCHS] += 1.0f;
    scalars[B200_SC_ADAM_STEP] += 1.0f;
    double lr = (double)scalars[B200_SC_LR];
    if (kl > desired_kl * 2.0f) lr = fmax((double)lr_min, lr / (double)lr_factor);
    else if (kl < desired_kl / 2.0f) lr = fmin((double)lr_max, lr * (double)lr_factor);
    scalars[B200_SC_LR] = (float)lr;
}

// clip + torch.optim.Adam step (default single-tensor math: lerp for m, mul/addcmul for v, bias corrections in fp64).  The step's
// scalars (clip coefficient, lr / bias correction 1, sqrt of bias correction 2: two fp64 pow and two fp64 sqrt) are evaluated by one
// thread per block - every thread used to evaluate them for itself, most of the kernel's 7.5 us - and the epoch's bookkeeping
// (post_apply_body: loss scalars, Adam step, KL-adaptive lr) runs in the LAST block to finish, i.e. after every block has read lr and the
// step (it used to be a launch of its own).
__global__ void __launch_bounds__(256) k_adam(float* __restrict__ params, float* __restrict__ grads, float* __restrict__ m1,
                                              float* __restrict__ m2, int n, float* __restrict__ scalars,
                                              const double* __restrict__ dstats, float inv_world, float max_norm,
                                              float beta1, float beta2, float eps, unsigned* __restrict__ done, float desired_kl,
                                              float lr_min, float lr_max, float lr_factor) {
    __shared__ float s_c[3];
    if (threadIdx.x == 0) {
        const float total_norm = (float)sqrt(dstats[DS_GRAD_SQ]);
        s_c[0] = fminf(__fdiv_rn(max_norm, __fadd_rn(total_norm, 1.0e-6f)), 1.0f);
        const double step = (double)scalars[B200_SC_ADAM_STEP] + 1.0;
        const double lr = (double)scalars[B200_SC_LR];
        const double bc1 = 1.0 - pow((double)beta1, step), bc2 = 1.0 - pow((double)beta2, step);
        s_c[1] = (float)(lr / bc1);
        s_c[2] = (float)sqrt(bc2);
    }
    __syncthreads();
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) {
        const float coef = s_c[0], step_size = s_c[1], bc2_sqrt = s_c[2];
        const float g = __fmul_rn(__fmul_rn(grads[i], inv_world), coef);
        grads[i] = g;  // the reference leaves the clipped gradient in .grad
        float m = m1[i], v = m2[i];
        m = __fadd_rn(m, __fmul_rn(__fsub_rn(g, m), 1.0f - beta1));
        v = __fadd_rn(__fmul_rn(v, beta2), __fmul_rn(__fmul_rn(g, g), 1.0f - beta2));
        m1[i] = m;
        m2[i] = v;
        const float denom = __fadd_rn(__fdiv_rn(sqrtf(v), bc2_sqrt), eps);
        params[i] = __fsub_rn(params[i], __fmul_rn(step_size, __fdiv_rn(m, denom)));
    }
    __syncthreads();   // every thread of this block has read the step's scalars
    if (threadIdx.x == 0) {
        __threadfence();
        if (atomicAdd(done, 1u) == gridDim.x - 1) {
            *done = 0u;
            post_apply_body(scalars, dstats, desired_kl, lr_min, lr_max, lr_factor);
        }
    }
}

// =====================================================================================================================
// measurement hooks: launch counter (bench.py "gpu_launches") and CUDA-event timing of every GEMM launch on the
// launching stream (bench.py "roofline": algorithmic FLOPs / measured duration of the dominant kernel)
// =====================================================================================================================
namespace b200 {
long long g_launches = 0;
}
enum { PK_MMA_SYNC = 0, PK_TC_ROW = 1, PK_TC_WGRAD = 2, PK_CHAIN_FWD = 3, PK_CHAIN_BWD = 4, PK_COUNT = 5 };  // kernel families timed by b200_profile_gemm
struct GemmProfile {
    bool on = false;
    int used = 0;
    static const int kMax = 8192;
    cudaEvent_t* ev = nullptr;  // 2 * kMax
    unsigned char* kind = nullptr;
    double flops[PK_COUNT] = {0.0, 0.0, 0.0, 0.0, 0.0};
    double bytes[PK_COUNT] = {0.0, 0.0, 0.0, 0.0, 0.0};   // HBM bytes of the launch's traffic model (operands read once + results written once)
} g_prof;

static bool g_tc_pair = getenv("B200_TC_PAIR") ? atoi(getenv("B200_TC_PAIR")) != 0 : false;   // cta_group::2 GEMMs (b200_tc_set_pair)
static int g_chain_exact_actor = getenv("B200_CHAIN_DEBUG") ? atoi(getenv("B200_CHAIN_DEBUG")) : 0;   // debug: 1 = single accumulator for the 128-wide layers
static bool g_chain_pair = getenv("B200_CHAIN_PAIR") ? atoi(getenv("B200_CHAIN_PAIR")) != 0 : false;   // chains on CTA pairs (cta_group::2)
static bool g_chain = getenv("B200_CHAIN") ? atoi(getenv("B200_CHAIN")) != 0 : true;           // fused layer chains (mlp_chain.cuh); 0 = layer-by-layer GEMMs
// b200_policy_act: rows from which the actor's hidden layers run on the tensor-core forward chain instead of the FP32-FMA kernel
// (128-row tcgen05 tiles: 4096 rows are 32 CTAs of one tile each; the FMA kernel needs ~46 us at 4096 rows and doubles at 8192)
static int g_policy_tc_min_rows = getenv("B200_POLICY_TC_MIN_ROWS") ? atoi(getenv("B200_POLICY_TC_MIN_ROWS")) : 2048;
static bool g_h2_chain = getenv("B200_H2") ? atoi(getenv("B200_H2")) != 0 : true;   // hidden-layer GEMMs on the h2 operand format (h2.cuh, mlp_chain_h2.cuh); 0 = 3xTF32 kernels
// epilogue warp groups of the h2 chains (1 or 2; measured equal within noise)
// h2 chains: weight multicast in clusters of two CTAs (halves the L2 -> SM weight stream; measured: no gain, the kernels are bound by the per-k-block
// epilogue latency, not by the fill - off by default)
static bool g_h2_mc = getenv("B200_H2_MC") ? atoi(getenv("B200_H2_MC")) != 0 : false;
static int g_h2_groups_fwd = getenv("B200_H2_GROUPS_FWD") ? atoi(getenv("B200_H2_GROUPS_FWD")) : 1;
static int g_h2_groups_bwd = getenv("B200_H2_GROUPS_BWD") ? atoi(getenv("B200_H2_GROUPS_BWD")) : 1;
static int g_tl_slot = 0;   // timeline slot of the next tcgen05 launch (debug builds)
static int tl_next() { const int s = g_tl_slot; g_tl_slot = (g_tl_slot + 1) % 40; return s; }
static void prof_begin(cudaStream_t st, double flops, int kind = PK_MMA_SYNC, double bytes = 0.0) {
    if (!g_prof.on || g_prof.used >= GemmProfile::kMax) return;
    cudaEventRecord(g_prof.ev[2 * g_prof.used], st);
    g_prof.kind[g_prof.used] = (unsigned char)kind;
    g_prof.flops[kind] += flops;
    g_prof.bytes[kind] += bytes;
}
static void prof_end(cudaStream_t st) {
    if (!g_prof.on || g_prof.used >= GemmProfile::kMax) return;
    cudaEventRecord(g_prof.ev[2 * g_prof.used + 1], st);
    g_prof.used += 1;
}

#define CU_TRY(expr)                                                   \
    do {                                                               \
        cudaError_t _e = (expr);                                       \
        if (_e != cudaSuccess) return set_cuda_error(_e, #expr);       \
    } while (0)

// ---- reduction of the k_tc_wgrad partial tiles: dW[row, col] += sum over parts, all six weight matrices in one launch ----
struct WgradJob {
    const float* P;   // [tiles_y * tiles_z][parts][128][bn]
    float* D;         // [n_out, ldd]
    int parts, bn, tiles_y, tiles_z, n_out, k_valid, ldd;
    int first;        // first linear output element of this job in the launch
};
struct WgradJobs {
    WgradJob job[6];
    int count, total;
};
#define WR_SLICES 8
__global__ void __launch_bounds__(256) k_wgrad_reduce(const WgradJobs J) {
    const int idx = blockIdx.x * 256 + threadIdx.x;
    if (idx >= J.total) return;
    int j = 0;
#pragma unroll
    for (int t = 1; t < 6; ++t)
        if (t < J.count && idx >= J.job[t].first) j = t;
    const WgradJob& w = J.job[j];
    const int e = idx - w.first;
    const int row = e / w.k_valid, col = e - row * w.k_valid;
    const int ty = row >> 7, tz = col / w.bn;
    const float* src = w.P + ((size_t)(tz * w.tiles_y + ty) * w.parts * 128 + (row & 127)) * w.bn + (col - tz * w.bn);
    const size_t stride = (size_t)128 * w.bn;
    const int per = (w.parts + WR_SLICES - 1) / WR_SLICES;
    const int p0 = blockIdx.y * per, p1 = min(w.parts, p0 + per);
    float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
    int pp = p0;
    for (; pp + 8 <= p1; pp += 8) {
        const float v0 = src[(size_t)pp * stride], v1 = src[(size_t)(pp + 1) * stride], v2 = src[(size_t)(pp + 2) * stride],
                    v3 = src[(size_t)(pp + 3) * stride], v4 = src[(size_t)(pp + 4) * stride], v5 = src[(size_t)(pp + 5) * stride],
                    v6 = src[(size_t)(pp + 6) * stride], v7 = src[(size_t)(pp + 7) * stride];
        s0 += v0 + v4; s1 += v1 + v5; s2 += v2 + v6; s3 += v3 + v7;
    }
    for (; pp < p1; ++pp) s0 += src[(size_t)pp * stride];
    if (p1 > p0) atomicAdd(w.D + (size_t)row * w.ldd + col, (s0 + s1) + (s2 + s3));
}

// ---- tcgen05 layers (gemm_tc.cuh) over pre-split operands ----------------------------------------------------------
#define TC_MAP(var, ptr, rows, cols, ld, box, kmajor)                                                      \
    const CUtensorMap* var = p->maps->get(ptr, rows, cols, ld, box, kmajor);                               \
    if (!var) return set_error(B200_ERR_CUDA, p->maps->error ? p->maps->error : "tensor map creation failed")

// Y [n, n_out] = ELU(X [n, k] W(h,l) [n_out, k_pad]^T + b)
static int tc_fwd(const B200Ppo* p, const float* X, int k, int ldx, const float* Wh, const float* Wl, int k_pad, const float* b,
                  float* Y, int n, int n_out, cudaStream_t st, bool accurate = false) {
    const int bn = (n_out >= 256 && !accurate) ? 256 : 128;
    const bool pair = !accurate && g_tc_pair;   // CTA pairs: each CTA stages half of the weight tile (box = bn / 2 rows)
    TC_MAP(mA, X, n, ldx, ldx, tc::BM, true);   // all ldx columns are stored (zero padded beyond k)
    TC_MAP(mBh, Wh, n_out, k_pad, k_pad, pair ? bn / 2 : bn, true);
    TC_MAP(mBl, Wl, n_out, k_pad, k_pad, pair ? bn / 2 : bn, true);
    tc::RowArgs g{};
    g.out = Y; g.bias = b; g.aux = nullptr; g.colsum = nullptr;
    g.M = n; g.Nout = n_out; g.K = k_pad; g.ldo = n_out; g.tl_slot = tl_next();
    prof_begin(st, 2.0 * n * (double)n_out * k, PK_TC_ROW, 4.0 * n * ((double)ldx + n_out));   // X read, Y written
    // <BN, NA raw-A stages, NB weight stages, epilogue, accumulators, CTA pair>: every configuration fills the 227 KB of shared memory
    const cudaError_t e = accurate    ? tc::launch_rowmajor<128, 5, 3, tc::EPI_FWD, 2>(mA, mBh, mBl, g, p->num_sms, st)
                          : pair ? ((bn == 256) ? tc::launch_rowmajor<256, 5, 3, tc::EPI_FWD, 1, 1>(mA, mBh, mBl, g, p->num_sms, st)
                                                : tc::launch_rowmajor<128, 8, 3, tc::EPI_FWD, 1, 1>(mA, mBh, mBl, g, p->num_sms, st))
                          : (bn == 256) ? tc::launch_rowmajor<256, 4, 2, tc::EPI_FWD>(mA, mBh, mBl, g, p->num_sms, st)
                                        : tc::launch_rowmajor<128, 5, 3, tc::EPI_FWD>(mA, mBh, mBl, g, p->num_sms, st);
    prof_end(st);
    g_launches += 1;
    if (e != cudaSuccess) return set_cuda_error(e, "k_tc_rowmajor<fwd>");
    return B200_OK;
}
// dX [n, k_in] = (dY [n, n_out] WT(h,l) [k_in, n_out]^T) * ELU'(H [n, k_in])
static int tc_dgrad(const B200Ppo* p, const float* dY, int n_out, const float* WTh, const float* WTl, int k_in, const float* H,
                    float* dX, float* colsum, int n, cudaStream_t st) {
    const int bn = (k_in >= 256) ? 256 : 128;
    const bool pair = g_tc_pair;
    TC_MAP(mA, dY, n, n_out, n_out, tc::BM, true);
    TC_MAP(mBh, WTh, k_in, n_out, n_out, pair ? bn / 2 : bn, true);
    TC_MAP(mBl, WTl, k_in, n_out, n_out, pair ? bn / 2 : bn, true);
    tc::RowArgs g{};
    g.out = dX; g.bias = nullptr; g.aux = H; g.colsum = colsum;
    g.M = n; g.Nout = k_in; g.K = n_out; g.ldo = k_in; g.tl_slot = tl_next();
    prof_begin(st, 2.0 * n * (double)n_out * k_in, PK_TC_ROW, 4.0 * n * ((double)n_out + 2.0 * k_in));   // dY, H read, dX written
    const cudaError_t e = pair ? ((bn == 256) ? tc::launch_rowmajor<256, 5, 3, tc::EPI_DGRAD, 1, 1>(mA, mBh, mBl, g, p->num_sms, st)
                                              : tc::launch_rowmajor<128, 8, 3, tc::EPI_DGRAD, 1, 1>(mA, mBh, mBl, g, p->num_sms, st))
                          : (bn == 256) ? tc::launch_rowmajor<256, 2, 2, tc::EPI_DGRAD>(mA, mBh, mBl, g, p->num_sms, st)
                                        : tc::launch_rowmajor<128, 3, 3, tc::EPI_DGRAD>(mA, mBh, mBl, g, p->num_sms, st);
    prof_end(st);
    g_launches += 1;
    if (e != cudaSuccess) return set_cuda_error(e, "k_tc_rowmajor<dgrad>");
    return B200_OK;
}
// dW [n_out, k_valid] += dY [n, n_out]^T X [n, k_cols]   (k_pad = 64 / 128 / 256 = tile width along k): launches the partial-tile
// GEMM into workspace region `slot` and appends the reduction job; wgrad_reduce() then sums all pending jobs in one launch
static int tc_wgrad(const B200Ppo* p, WgradJobs& jobs, int slot, const float* dY, int n_out, const float* X, int k_cols, int k_pad,
                    int k_valid, float* dW, int n, cudaStream_t st) {
    const int wbk = (k_pad == 256) ? 16 : 32;   // rows per pipeline stage (= TMA box rows)
    const int bn = k_pad >= 256 ? 256 : k_pad;
    const CUtensorMap* mY = p->maps->get3(dY, n, n_out, n_out, wbk, tc::BM / 32);
    const CUtensorMap* mX = p->maps->get3(X, n, k_cols, k_cols, wbk, bn / 32);
    if (!mY || !mX) return set_error(B200_ERR_CUDA, p->maps->error ? p->maps->error : "tensor map creation failed");
    tc::WgradArgs g{};
    g.P = p->ws + p->w.WP[slot]; g.M = n; g.Nout = n_out; g.Kin = k_valid; g.tl_slot = tl_next();
    const int tiles_y = (n_out + tc::BM - 1) / tc::BM, tiles_z = (k_pad + bn - 1) / bn;
    int parts;
    {
        // one wave: the CTAs of an output tile split the rows evenly (multiples of the 32-row k-block)
        const int tiles = tiles_y * tiles_z;
        const int sms = p->num_sms < WP_MAX_CTAS ? p->num_sms : WP_MAX_CTAS;
        const int per_tile = sms / tiles > 0 ? sms / tiles : 1;
        const int rows = (n + per_tile - 1) / per_tile;
        g.chunk = ((rows + 31) / 32) * 32;
        parts = (n + g.chunk - 1) / g.chunk;
    }
    WgradJob& j = jobs.job[jobs.count];
    j.P = g.P; j.D = dW; j.parts = parts; j.bn = bn; j.tiles_y = tiles_y; j.tiles_z = tiles_z; j.n_out = n_out; j.k_valid = k_valid;
    j.ldd = k_valid; j.first = jobs.total;
    jobs.count += 1;
    jobs.total += n_out * k_valid;
    prof_begin(st, 2.0 * n * (double)n_out * k_valid, PK_TC_WGRAD, 4.0 * n * ((double)n_out + k_cols));   // dY, X read
    cudaError_t e;
    if (k_pad == 256) e = tc::launch_wgrad<256, 4, 16>(mY, mX, g, k_pad, st);
    else if (k_pad == 128) e = tc::launch_wgrad<128, 3, 32>(mY, mX, g, k_pad, st);
    else e = tc::launch_wgrad<64, 4, 32>(mY, mX, g, k_pad, st);
    prof_end(st);
    g_launches += 1;
    if (e != cudaSuccess) return set_cuda_error(e, "k_tc_wgrad");
    return B200_OK;
}
static int wgrad_reduce(const WgradJobs& jobs, cudaStream_t st) {
    k_wgrad_reduce<<<dim3((jobs.total + 255) / 256, WR_SLICES), 256, 0, st>>>(jobs);
    g_launches += 1;
    return launch_status("k_wgrad_reduce");
}
// one small launch instead of two memset nodes (back-to-back memsets cost ~5 us of idle time each inside a graph)
__global__ void k_zero2(float* __restrict__ a, int na, double* __restrict__ b, int nb) {
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < na; i += gridDim.x * blockDim.x) a[i] = 0.0f;
    if (b) for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < nb; i += gridDim.x * blockDim.x) b[i] = 0.0;
}
// split copies of the six hidden-layer weight matrices (they change every epoch); `also_zero` (nullable): the epoch's dstats
static int weight_prep(const B200Ppo* p, cudaStream_t st, double* also_zero = nullptr) {
    float* ws = p->ws;
    const Workspace& w = p->w;
    struct Job { int pi, rows, cols, pad; size_t h, l, th, tl; bool tr; };
    const Job jobs[6] = {
        {P_CW0, 256, 61, 64, w.Wc0h, w.Wc0l, 0, 0, false},        {P_CW1, 256, 256, 256, w.Wc1h, w.Wc1l, w.Wc1Th, w.Wc1Tl, true},
        {P_CW2, 128, 256, 256, w.Wc2h, w.Wc2l, w.Wc2Th, w.Wc2Tl, true}, {P_AW0, 256, 47, 64, w.Wa0h, w.Wa0l, 0, 0, false},
        {P_AW1, 128, 256, 256, w.Wa1h, w.Wa1l, w.Wa1Th, w.Wa1Tl, true}, {P_AW2, 128, 128, 128, w.Wa2h, w.Wa2l, w.Wa2Th, w.Wa2Tl, true}};
    WeightPrepJobs wj;
    int max_total = 0;
    // column abs-sums of the four matrices an input gradient passes through: [0] critic.4, [1] critic.2, [2] actor.4, [3] actor.2
    static const int cab_slot[6] = {-1, 1, 0, -1, 3, 2};
    k_zero2<<<2, 256, 0, st>>>(ws + w.ZR, 64 + 4 * 256, also_zero, DS_COUNT);
    g_launches += 1;
    for (int i = 0; i < 6; ++i) {
        const Job& j = jobs[i];
        wj.j[i] = WeightPrepJob{p->P(j.pi), ws + j.h, ws + j.l, j.tr ? ws + j.th : nullptr, j.tr ? ws + j.tl : nullptr, j.rows, j.cols, j.pad,
                                cab_slot[i] >= 0 ? ws + w.ZR + 64 + cab_slot[i] * 256 : nullptr};
        max_total = j.rows * j.pad > max_total ? j.rows * j.pad : max_total;
    }
    wj.h2 = g_h2_chain ? 1 : 0;
    wj.bias[0] = p->P(P_CB0); wj.bias[1] = p->P(P_CB1); wj.bias[2] = p->P(P_AB0); wj.bias[3] = p->P(P_AB1);
    wj.bias_n[0] = 256; wj.bias_n[1] = 256; wj.bias_n[2] = 256; wj.bias_n[3] = 128;
    wj.bias_scaled = ws + w.BS;
    k_weight_prep<<<dim3((max_total + 255) / 256, g_h2_chain ? 7 : 6), 256, 0, st>>>(wj);
    g_launches += 1;
    return launch_status("k_weight_prep");
}
// full-batch forward passes over the M = T*N stored samples (activations kept in fp32 for the backward pass)
// The ACTOR forward uses the 4-accumulator variant with the exact expm1f: log-prob sensitivity to mu is 1/sigma ~ 7.4 per
// unit (sigma = e^-2), so mu wants short accumulation chains (TMEM adds truncate) - see gemm_tc.cuh.
// ---- fused layer chains (mlp_chain.cuh): the three hidden layers of a net in one persistent kernel per direction --------------
struct ChainNetPtrs {          // one net's operands of a forward launch
    const float *Xh, *Xl;      // [rows, 64] pre-split input
    const float *W1h, *W1l, *W2h, *W2l, *W3h, *W3l, *b1, *b2, *b3;
    float *H1, *H2, *H3;
    int rows, n2, exact, k_valid;
};
static int chain_fill_fwd(const B200Ppo* p, chain::FwdNet& N, const ChainNetPtrs& c) {
    N.rows = c.rows; N.n2 = c.n2; N.exact = c.exact; N.pad_ = 0;
    if (c.rows <= 0) { N.rows = 0; return B200_OK; }
    TC_MAP(mXh, c.Xh, c.rows, 64, 64, tc::BM, true);
    TC_MAP(mXl, c.Xl, c.rows, 64, 64, tc::BM, true);
    const int pd = g_chain_pair ? 2 : 1;   // CTA pairs: each CTA stages half of the rows of a weight k-block
    TC_MAP(mW1h, c.W1h, 256, 64, 64, 256 / pd, true);
    TC_MAP(mW1l, c.W1l, 256, 64, 64, 256 / pd, true);
    TC_MAP(mW2h, c.W2h, c.n2, 256, 256, c.n2 / pd, true);
    TC_MAP(mW2l, c.W2l, c.n2, 256, 256, c.n2 / pd, true);
    TC_MAP(mW3h, c.W3h, 128, c.n2, c.n2, 128 / pd, true);
    TC_MAP(mW3l, c.W3l, 128, c.n2, c.n2, 128 / pd, true);
    N.mXh = *mXh; N.mXl = *mXl; N.mW1h = *mW1h; N.mW1l = *mW1l; N.mW2h = *mW2h; N.mW2l = *mW2l; N.mW3h = *mW3h; N.mW3l = *mW3l;
    N.b1 = c.b1; N.b2 = c.b2; N.b3 = c.b3; N.H1 = c.H1; N.H2 = c.H2; N.H3 = c.H3;
    return B200_OK;
}
static ChainNetPtrs critic_ptrs(const B200Ppo* p, int rows) {
    float* ws = p->ws; const Workspace& w = p->w;
    return ChainNetPtrs{ws + w.Xch, ws + w.Xcl, ws + w.Wc0h, ws + w.Wc0l, ws + w.Wc1h, ws + w.Wc1l, ws + w.Wc2h, ws + w.Wc2l,
                        p->P(P_CB0), p->P(P_CB1), p->P(P_CB2), ws + w.C1, ws + w.C2, ws + w.C3, rows, 256, g_chain_exact_actor, 61};
}
static ChainNetPtrs actor_ptrs(const B200Ppo* p, int rows) {
    float* ws = p->ws; const Workspace& w = p->w;
    return ChainNetPtrs{ws + w.Xah, ws + w.Xal, ws + w.Wa0h, ws + w.Wa0l, ws + w.Wa1h, ws + w.Wa1l, ws + w.Wa2h, ws + w.Wa2l,
                        p->P(P_AB0), p->P(P_AB1), p->P(P_AB2), ws + w.A1, ws + w.A2, ws + w.A3, rows, 128, g_chain_exact_actor, 47};
}
// one persistent CTA per SM, or (pair) one cluster of two CTAs per TPC
template <typename Params>
static cudaError_t chain_launch(void (*kernel)(const Params), const Params& P, int rows0, int rows1, int threads, int smem, bool pair,
                                int num_sms, cudaStream_t st) {
    const int tm = pair ? 2 * tc::BM : tc::BM;
    const int tiles = (rows0 + tm - 1) / tm + (rows1 + tm - 1) / tm;
    cudaLaunchConfig_t cfg = {};
    cudaLaunchAttribute attr[1];
    if (pair) {
        const int pairs = tiles < num_sms / 2 ? tiles : num_sms / 2;
        cfg.gridDim = dim3(2 * pairs);
        attr[0].id = cudaLaunchAttributeClusterDimension;
        attr[0].val.clusterDim.x = 2;
        attr[0].val.clusterDim.y = 1;
        attr[0].val.clusterDim.z = 1;
        cfg.attrs = attr;
        cfg.numAttrs = 1;
    } else {
        cfg.gridDim = dim3(tiles < num_sms ? tiles : num_sms);
    }
    cfg.blockDim = dim3(threads);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    return cudaLaunchKernelEx(&cfg, kernel, P);
}

// h2 chains: one persistent CTA per SM; `pair`: clusters of two CTAs that share the weight stream (256-row work tiles)
template <typename Params>
static cudaError_t h2_chain_launch(void (*kernel)(const Params), const Params& P, int rows0, int rows1, int threads, int smem, bool pair,
                                   int num_sms, cudaStream_t st) {
    return chain_launch(kernel, P, rows0, rows1, threads, smem, pair, num_sms, st);
}
// midpoint compensation of a TMEM accumulator built by n truncating adds (mlp_chain_h2.cuh, FwdNet::gain): 1 + n 2^-26, scaled by
// B200_H2_MIDPOINT (default 1; 0 switches the compensation off)
static float midpoint_gain(int n_adds, int which = 0) {
    static const float k = getenv("B200_H2_MIDPOINT") ? (float)atof(getenv("B200_H2_MIDPOINT")) : 1.0f;
    static const float kx[4] = {getenv("B200_H2_MIDPOINT_L1") ? (float)atof(getenv("B200_H2_MIDPOINT_L1")) : 1.0f,
                                getenv("B200_H2_MIDPOINT_L2") ? (float)atof(getenv("B200_H2_MIDPOINT_L2")) : 1.0f,
                                getenv("B200_H2_MIDPOINT_L3") ? (float)atof(getenv("B200_H2_MIDPOINT_L3")) : 1.0f,
                                getenv("B200_H2_MIDPOINT_BWD") ? (float)atof(getenv("B200_H2_MIDPOINT_BWD")) : 1.0f};   // measurement knobs
    return 1.0f + k * kx[which & 3] * (float)n_adds * 1.490116119384765625e-8f;
}
// the same launch on the h2 operand format (mlp_chain_h2.cuh): Xh = input words, W*h / W*l = B' / B'' words, H1 / H2 receive words
static int chain_forward_h2(B200Ppo* p, const ChainNetPtrs& c0, const ChainNetPtrs& c1, cudaStream_t st) {
    chain2::FwdParams P;
    memset(&P, 0, sizeof(P));
    const ChainNetPtrs* cs[2] = {&c0, &c1};
    double fl = 0.0, by = 0.0;
    for (int i = 0; i < 2; ++i) {
        const ChainNetPtrs& c = *cs[i];
        chain2::FwdNet& N = P.net[i];
        N.rows = c.rows > 0 ? c.rows : 0; N.n2 = c.n2; N.pad0_ = 0; N.pad_ = 0;
        if (c.rows <= 0) continue;
        const int pd = g_h2_mc ? 2 : 1;   // weight multicast in CTA pairs: each CTA loads half of the rows of a weight k-block
        TC_MAP(mX, c.Xh, c.rows, 64, 64, tc::BM, true);
        TC_MAP(mW1a, c.W1h, 256, 64, 64, 256 / pd, true);
        TC_MAP(mW1b, c.W1l, 256, 64, 64, 256 / pd, true);
        TC_MAP(mW2a, c.W2h, c.n2, 256, 256, (c.n2 == 256 ? 256 : 128) / pd, true);
        TC_MAP(mW2b, c.W2l, c.n2, 256, 256, (c.n2 == 256 ? 256 : 128) / pd, true);
        TC_MAP(mW3a, c.W3h, 128, c.n2, c.n2, 128 / pd, true);
        TC_MAP(mW3b, c.W3l, 128, c.n2, c.n2, 128 / pd, true);
        TC_MAP(mH1, c.H1, c.rows, 256, 256, tc::BM, true);
        TC_MAP(mH2, c.H2, c.rows, c.n2, c.n2, tc::BM, true);
        TC_MAP(mH3, c.H3, c.rows, 128, 128, tc::BM, true);
        N.mX = *mX; N.mW1a = *mW1a; N.mW1b = *mW1b; N.mW2a = *mW2a; N.mW2b = *mW2b; N.mW3a = *mW3a; N.mW3b = *mW3b;
        N.mH1 = *mH1; N.mH2 = *mH2; N.mH3 = *mH3;
        N.b1s = p->ws + p->w.BS + (i == 0 ? 0 : 512); N.b2s = N.b1s + 256; N.b3 = c.b3;   // (net 0 = critic, net 1 = actor)
        // accumulating MMAs per layer with the dominant terms present: layer 1 issues its 8 cross-term MMAs first (into a small
        // accumulator) and 8 main ones; layer 2 is 8 k-blocks x 4 k-steps (x 2 when the 256-wide layer issues cross and main terms
        // into the same columns); layer 3 is n2 / 32 k-blocks x 4 on stacked main | cross columns
        N.gain[0] = midpoint_gain(8, 0); N.gain[1] = midpoint_gain(c.n2 == 256 ? 64 : 32, 1); N.gain[2] = midpoint_gain(c.n2 / 32 * 4, 2); N.gain[3] = 1.0f;
        fl += 2.0 * c.rows * ((double)c.k_valid * 256 + 256.0 * c.n2 + (double)c.n2 * 128);
        by += 4.0 * c.rows * (64 + 256 + c.n2 + 128);   // X words read; h1, h2 words and h3 written
    }
    const int tiles = (P.net[0].rows + tc::BM - 1) / tc::BM + (P.net[1].rows + tc::BM - 1) / tc::BM;
    if (tiles <= 0) return B200_OK;
    static unsigned long long configured[4] = {0, 0, 0, 0};
    CU_TRY(ensure_dynamic_smem(chain2::k_mlp_fwd_h2<1, 0>, chain2::F_SMEM, configured[0]));
    CU_TRY(ensure_dynamic_smem(chain2::k_mlp_fwd_h2<2, 0>, chain2::F_SMEM, configured[1]));
    CU_TRY(ensure_dynamic_smem(chain2::k_mlp_fwd_h2<1, 1>, chain2::F_SMEM, configured[2]));
    CU_TRY(ensure_dynamic_smem(chain2::k_mlp_fwd_h2<2, 1>, chain2::F_SMEM, configured[3]));
    prof_begin(st, fl, PK_CHAIN_FWD, by);
    const bool g2 = g_h2_groups_fwd != 1;
    const cudaError_t le = g_h2_mc ? h2_chain_launch(g2 ? chain2::k_mlp_fwd_h2<2, 1> : chain2::k_mlp_fwd_h2<1, 1>, P, P.net[0].rows, P.net[1].rows, chain2::F_THREADS, chain2::F_SMEM, true, p->num_sms, st)
                                   : h2_chain_launch(g2 ? chain2::k_mlp_fwd_h2<2, 0> : chain2::k_mlp_fwd_h2<1, 0>, P, P.net[0].rows, P.net[1].rows, chain2::F_THREADS, chain2::F_SMEM, false, p->num_sms, st);
    prof_end(st);
    g_launches += 1;
    if (le != cudaSuccess) return set_cuda_error(le, "k_mlp_fwd_h2");
    return launch_status("k_mlp_fwd_h2");
}
static int chain_forward(B200Ppo* p, const ChainNetPtrs& c0, const ChainNetPtrs& c1, cudaStream_t st) {
    if (g_h2_chain) return chain_forward_h2(p, c0, c1, st);
    chain::FwdParams P;
    memset(&P, 0, sizeof(P));
    int rc;
    if ((rc = chain_fill_fwd(p, P.net[0], c0)) != B200_OK) return rc;
    if ((rc = chain_fill_fwd(p, P.net[1], c1)) != B200_OK) return rc;
    const int tiles = (P.net[0].rows + tc::BM - 1) / tc::BM + (P.net[1].rows + tc::BM - 1) / tc::BM;
    if (tiles <= 0) return B200_OK;
    if (!p->chain_fwd_configured) {
        CUDA_TRY(cudaFuncSetAttribute(chain::k_mlp_fwd<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, chain::F_SMEM));
        CUDA_TRY(cudaFuncSetAttribute(chain::k_mlp_fwd<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, chain::F_SMEM));
        p->chain_fwd_configured = true;
    }
    double fl = 0.0, by = 0.0;
    const ChainNetPtrs* cs[2] = {&c0, &c1};
    for (int i = 0; i < 2; ++i) {
        const double r = cs[i]->rows > 0 ? cs[i]->rows : 0;
        fl += 2.0 * r * ((double)cs[i]->k_valid * 256 + 256.0 * cs[i]->n2 + (double)cs[i]->n2 * 128);
        by += 4.0 * r * (2 * 64 + 256 + cs[i]->n2 + 128);   // X hi + lo read; h1, h2, h3 written
    }
    prof_begin(st, fl, PK_CHAIN_FWD, by);
    const cudaError_t le = g_chain_pair ? chain_launch(chain::k_mlp_fwd<1>, P, P.net[0].rows, P.net[1].rows, chain::F_THREADS, chain::F_SMEM, true, p->num_sms, st)
                                        : chain_launch(chain::k_mlp_fwd<0>, P, P.net[0].rows, P.net[1].rows, chain::F_THREADS, chain::F_SMEM, false, p->num_sms, st);
    prof_end(st);
    g_launches += 1;
    if (le != cudaSuccess) return set_cuda_error(le, "k_mlp_fwd");
    return launch_status("k_mlp_fwd");
}
// both nets' hidden-layer input gradients + bias gradients: dz3 (from the head kernels) -> dz2, dz1
static int chain_backward_h2(B200Ppo* p, int M, bool critic, bool actor, cudaStream_t st) {
    float* ws = p->ws; const Workspace& w = p->w;
    chain2::BwdParams P;
    memset(&P, 0, sizeof(P));
    double fl = 0.0, by = 0.0;
    for (int i = 0; i < 2; ++i) {
        chain2::BwdNet& N = P.net[i];
        N.n2 = i == 0 ? 256 : 128;
        N.rows = (i == 0 ? critic : actor) ? M : 0;
        // dz2 = dz3 W3: 4 k-blocks x 4 k-steps (x 2 MMAs into the same columns for the 256-wide critic layer); dz1 = dz2 W2: n2 / 32
        // k-blocks x 4 k-steps x 2 (the 256 outputs take the cross and the main MMA in the same columns)
        N.gain[0] = midpoint_gain(N.n2 == 256 ? 32 : 16, 3); N.gain[1] = midpoint_gain(N.n2 / 32 * 4 * 2, 3);
        if (N.rows <= 0) continue;
        TC_MAP(mZ3, ws + (i == 0 ? w.GC3 : w.GA3), M, 128, 128, tc::BM, true);
        TC_MAP(mH2, ws + (i == 0 ? w.C2 : w.A2), M, N.n2, N.n2, tc::BM, true);
        TC_MAP(mH1, ws + (i == 0 ? w.C1 : w.A1), M, 256, 256, tc::BM, true);
        const int pd = g_h2_mc ? 2 : 1;
        TC_MAP(mW3Ta, ws + (i == 0 ? w.Wc2Th : w.Wa2Th), N.n2, 128, 128, (N.n2 == 256 ? 256 : 128) / pd, true);
        TC_MAP(mW3Tb, ws + (i == 0 ? w.Wc2Tl : w.Wa2Tl), N.n2, 128, 128, (N.n2 == 256 ? 256 : 128) / pd, true);
        TC_MAP(mW2Ta, ws + (i == 0 ? w.Wc1Th : w.Wa1Th), 256, N.n2, N.n2, 256 / pd, true);
        TC_MAP(mW2Tb, ws + (i == 0 ? w.Wc1Tl : w.Wa1Tl), 256, N.n2, N.n2, 256 / pd, true);
        TC_MAP(mDZ2, ws + (i == 0 ? w.GC2 : w.GA2), M, N.n2, N.n2, tc::BM, true);
        TC_MAP(mDZ1, ws + (i == 0 ? w.GC1 : w.GA1), M, 256, 256, tc::BM, true);
        N.mZ3 = *mZ3; N.mH2 = *mH2; N.mH1 = *mH1; N.mW3Ta = *mW3Ta; N.mW3Tb = *mW3Tb; N.mW2Ta = *mW2Ta; N.mW2Tb = *mW2Tb;
        N.mDZ2 = *mDZ2; N.mDZ1 = *mDZ1;
        N.db2 = p->G(i == 0 ? P_CB1 : P_AB1);
        N.db1 = p->G(i == 0 ? P_CB0 : P_AB0);
        N.isg = ws + w.SC + (i == 0 ? h2::SC_ISG_C : h2::SC_ISG_A);
        fl += 2.0 * M * (128.0 * N.n2 + (double)N.n2 * 256);
        by += 4.0 * M * (128.0 + 2.0 * N.n2 + 2.0 * 256);   // dz3, h2, h1 read; dz2, dz1 written
    }
    const int tiles = (P.net[0].rows + tc::BM - 1) / tc::BM + (P.net[1].rows + tc::BM - 1) / tc::BM;
    if (tiles <= 0) return B200_OK;
    static unsigned long long configured[4] = {0, 0, 0, 0};
    CU_TRY(ensure_dynamic_smem(chain2::k_mlp_bwd_h2<1, 0>, chain2::B_SMEM, configured[0]));
    CU_TRY(ensure_dynamic_smem(chain2::k_mlp_bwd_h2<2, 0>, chain2::B_SMEM, configured[1]));
    CU_TRY(ensure_dynamic_smem(chain2::k_mlp_bwd_h2<1, 1>, chain2::B_SMEM, configured[2]));
    CU_TRY(ensure_dynamic_smem(chain2::k_mlp_bwd_h2<2, 1>, chain2::B_SMEM, configured[3]));
    prof_begin(st, fl, PK_CHAIN_BWD, by);
    const bool g2 = g_h2_groups_bwd != 1;
    const cudaError_t le = g_h2_mc ? h2_chain_launch(g2 ? chain2::k_mlp_bwd_h2<2, 1> : chain2::k_mlp_bwd_h2<1, 1>, P, P.net[0].rows, P.net[1].rows, chain2::B_THREADS, chain2::B_SMEM, true, p->num_sms, st)
                                   : h2_chain_launch(g2 ? chain2::k_mlp_bwd_h2<2, 0> : chain2::k_mlp_bwd_h2<1, 0>, P, P.net[0].rows, P.net[1].rows, chain2::B_THREADS, chain2::B_SMEM, false, p->num_sms, st);
    prof_end(st);
    g_launches += 1;
    if (le != cudaSuccess) return set_cuda_error(le, "k_mlp_bwd_h2");
    return launch_status("k_mlp_bwd_h2");
}
static int chain_backward(B200Ppo* p, int M, bool critic, bool actor, cudaStream_t st) {
    if (g_h2_chain) return chain_backward_h2(p, M, critic, actor, st);
    float* ws = p->ws; const Workspace& w = p->w;
    chain::BwdParams P;
    memset(&P, 0, sizeof(P));
    double fl = 0.0, by = 0.0;
    for (int i = 0; i < 2; ++i) {
        chain::BwdNet& N = P.net[i];
        N.n2 = i == 0 ? 256 : 128;
        N.rows = (i == 0 ? critic : actor) ? M : 0;
        if (N.rows <= 0) continue;
        const float* Z3 = ws + (i == 0 ? w.GC3 : w.GA3);
        const float* H2 = ws + (i == 0 ? w.C2 : w.A2);
        const float* H1 = ws + (i == 0 ? w.C1 : w.A1);
        const float* W3Th = ws + (i == 0 ? w.Wc2Th : w.Wa2Th);
        const float* W3Tl = ws + (i == 0 ? w.Wc2Tl : w.Wa2Tl);
        const float* W2Th = ws + (i == 0 ? w.Wc1Th : w.Wa1Th);
        const float* W2Tl = ws + (i == 0 ? w.Wc1Tl : w.Wa1Tl);
        TC_MAP(mZ3, Z3, M, 128, 128, tc::BM, true);
        TC_MAP(mH2, H2, M, N.n2, N.n2, tc::BM, true);
        TC_MAP(mH1, H1, M, 256, 256, tc::BM, true);
        TC_MAP(mW3Th, W3Th, N.n2, 128, 128, g_chain_pair ? N.n2 / 2 : N.n2, true);
        TC_MAP(mW3Tl, W3Tl, N.n2, 128, 128, g_chain_pair ? N.n2 / 2 : N.n2, true);
        TC_MAP(mW2Th, W2Th, 256, N.n2, N.n2, g_chain_pair ? 64 : 256, true);   // pairs: 64 rows of each 128-column output half per CTA
        TC_MAP(mW2Tl, W2Tl, 256, N.n2, N.n2, g_chain_pair ? 64 : 256, true);
        N.mZ3 = *mZ3; N.mH2 = *mH2; N.mH1 = *mH1; N.mW3Th = *mW3Th; N.mW3Tl = *mW3Tl; N.mW2Th = *mW2Th; N.mW2Tl = *mW2Tl;
        N.DZ2 = ws + (i == 0 ? w.GC2 : w.GA2);
        N.DZ1 = ws + (i == 0 ? w.GC1 : w.GA1);
        N.db2 = p->G(i == 0 ? P_CB1 : P_AB1);
        N.db1 = p->G(i == 0 ? P_CB0 : P_AB0);
        fl += 2.0 * M * (128.0 * N.n2 + (double)N.n2 * 256);
        by += 4.0 * M * (128.0 + 2.0 * N.n2 + 2.0 * 256);   // dz3, h2, h1 read; dz2, dz1 written
    }
    const int tiles = (P.net[0].rows + tc::BM - 1) / tc::BM + (P.net[1].rows + tc::BM - 1) / tc::BM;
    if (tiles <= 0) return B200_OK;
    if (!p->chain_bwd_configured) {
        CUDA_TRY(cudaFuncSetAttribute(chain::k_mlp_bwd<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, chain::B_SMEM));
        CUDA_TRY(cudaFuncSetAttribute(chain::k_mlp_bwd<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, chain::B_SMEM));
        p->chain_bwd_configured = true;
    }
    prof_begin(st, fl, PK_CHAIN_BWD, by);
    const cudaError_t le = g_chain_pair ? chain_launch(chain::k_mlp_bwd<1>, P, P.net[0].rows, P.net[1].rows, chain::B_THREADS, chain::B_SMEM, true, p->num_sms, st)
                                        : chain_launch(chain::k_mlp_bwd<0>, P, P.net[0].rows, P.net[1].rows, chain::B_THREADS, chain::B_SMEM, false, p->num_sms, st);
    prof_end(st);
    g_launches += 1;
    if (le != cudaSuccess) return set_cuda_error(le, "k_mlp_bwd");
    return launch_status("k_mlp_bwd");
}

// ---- weight gradients on the h2 operand format (h2.cuh): all six hidden-layer matrices in one launch + one reduction ----------
struct WgH2Operand { const float* dY; const float* X; int n_out, k_cols, k_valid; int pi; bool actor; float sx; };
static int wgrad_h2(B200Ppo* p, const WgH2Operand* ops, int M, cudaStream_t st) {
    h2::WgParams P;
    h2::WgRedJobs R;
    memset(&P, 0, sizeof(P));
    memset(&R, 0, sizeof(R));
    // tiles (n-major, then k) and their cost per row
    double cost[h2::WG_MAX_TILES];
    int nt = 0, total = 0;
    double flops = 0.0, bytes = 0.0;
    for (int j = 0; j < 6; ++j) {
        const WgH2Operand& o = ops[j];
        const int kch = o.k_cols >= 128 ? 4 : 2;
        const CUtensorMap* mY = p->maps->get3(o.dY, M, o.n_out, o.n_out, h2::WG2_ROWS, 4, true);
        const CUtensorMap* mX = p->maps->get3(o.X, M, o.k_cols, o.k_cols, h2::WG2_ROWS, kch, true);
        if (!mY || !mX) return set_error(B200_ERR_CUDA, p->maps->error ? p->maps->error : "tensor map creation failed");
        P.mY[j] = *mY; P.mX[j] = *mX; P.kchunks[j] = kch;
        const int tn = o.n_out / 128, tk = (o.k_cols + 127) / 128;
        h2::WgRedJob& r = R.job[j];
        r.D = p->G(o.pi); r.isg = p->ws + p->w.SC + (o.actor ? h2::SC_ISG_A : h2::SC_ISG_C); r.inv_sx = 1.0f / o.sx;
        r.n_out = o.n_out; r.k_valid = o.k_valid; r.ldd = o.k_valid; r.tile0 = nt; r.tiles_k = tk; r.first = total;
        total += o.n_out * o.k_valid;
        for (int a = 0; a < tn; ++a)
            for (int b = 0; b < tk; ++b, ++nt) {
                if (nt >= h2::WG_MAX_TILES) return set_error(B200_ERR_ARG, "wgrad_h2: too many tiles");
                P.tile[nt].job = j; P.tile[nt].n0 = a * 128; P.tile[nt].k0 = b * 128;
                cost[nt] = 1.0;   // an MMA costs the same ~216 clk whether N is 256 or 128 (tools/micro/mma_rate.cu)
            }
        flops += 2.0 * M * (double)o.n_out * o.k_valid;
        bytes += 4.0 * M * ((double)o.n_out + o.k_cols);
    }
    // one wave: the smallest per-CTA budget B with sum ceil(M cost / B) <= CTAs
    const int ctas = p->num_sms < WP_MAX_CTAS ? p->num_sms : WP_MAX_CTAS;
    double lo = 1.0, hi = (double)M * 2.0;
    for (int it = 0; it < 60; ++it) {
        const double mid = 0.5 * (lo + hi);
        long long need = 0;
        for (int t = 0; t < nt; ++t) need += (long long)ceil(M * cost[t] / mid);
        if (need <= ctas) hi = mid; else lo = mid;
    }
    int first = 0;
    for (int t = 0; t < nt; ++t) {
        int parts = (int)ceil(M * cost[t] / hi);
        if (parts < 1) parts = 1;
        int chunk = ((M + parts - 1) / parts + 31) / 32 * 32;
        parts = (M + chunk - 1) / chunk;
        P.tile[t].first = first; P.tile[t].parts = parts; P.tile[t].chunk = chunk;
        first += parts;
    }
    if (first > WP_MAX_CTAS) return set_error(B200_ERR_ARG, "wgrad_h2: work split exceeds the partial-tile buffer");
    P.ntiles = nt; P.M = M; P.P = p->ws + p->w.WP2;
    memcpy(R.tile, P.tile, sizeof(P.tile));
    R.P = P.P; R.count = 6; R.total = total;
    static unsigned long long configured = 0;
    CU_TRY(ensure_dynamic_smem(h2::k_wgrad_h2, h2::WG2_SMEM, configured));
    prof_begin(st, flops, PK_TC_WGRAD, bytes);
    h2::k_wgrad_h2<<<first, h2::WG2_THREADS, h2::WG2_SMEM, st>>>(P);
    prof_end(st);
    h2::k_wgrad_h2_reduce<<<(total + 255) / 256, 256, 0, st>>>(R);
    g_launches += 2;
    return launch_status("k_wgrad_h2");
}

// forward heads on the bulk-copy-fed streaming kernels (heads.cuh)
template <int J>
static int head_forward(const B200Ppo* p, const float* H, const float* W, const float* b, int n, float* out, cudaStream_t st) {
    constexpr int SMEM = heads::HB2_STAGES * heads::HB2_ROWS * 512 + J * 512 + 64;
    static unsigned long long configured = 0;
    CU_TRY(ensure_dynamic_smem(heads::k_head_fwd_pipe<J>, SMEM, configured));
    const int tiles = (n + heads::HB2_ROWS - 1) / heads::HB2_ROWS, grid = tiles < 2 * p->num_sms ? tiles : 2 * p->num_sms;
    heads::k_head_fwd_pipe<J><<<grid, heads::HB2_THREADS, SMEM, st>>>(H, W, b, n, out);
    g_launches += 1;
    return launch_status("k_head_fwd_pipe");
}

static int actor_forward_tc(B200Ppo* p, int M, cudaStream_t st) {
    float* ws = p->ws;
    const Workspace& w = p->w;
    int rc;
    if (g_chain) {
        // the SAME head kernel as the epoch's forward (b200_ppo_epoch_a): the old distribution of b200_ppo_old_dist and the first epoch's mu
        // must agree to the bit, as they do in the reference (ratio == 1, KL == 0 in epoch 0) - a different summation order moved mu by
        // ~1e-7 of its scale and, amplified by (a - mu) / sigma^2, the probability ratios by up to 3e-4
        if ((rc = chain_forward(p, critic_ptrs(p, 0), actor_ptrs(p, M), st))) return rc;
        return head_forward<12>(p, ws + w.A3, p->P(P_AW3), p->P(P_AB3), M, ws + w.MU, st);
    }
    if ((rc = tc_fwd(p, ws + w.Xa, 47, 64, ws + w.Wa0h, ws + w.Wa0l, 64, p->P(P_AB0), ws + w.A1, M, 256, st, true))) return rc;
    if ((rc = tc_fwd(p, ws + w.A1, 256, 256, ws + w.Wa1h, ws + w.Wa1l, 256, p->P(P_AB1), ws + w.A2, M, 128, st, true))) return rc;
    if ((rc = tc_fwd(p, ws + w.A2, 128, 128, ws + w.Wa2h, ws + w.Wa2l, 128, p->P(P_AB2), ws + w.A3, M, 128, st, true))) return rc;
    k_actor_head<<<1184, 256, 0, st>>>(ws + w.A3, p->P(P_AW3), p->P(P_AB3), M, ws + w.MU);
    g_launches += 1;
    return launch_status("k_actor_head");
}
static int critic_forward_tc(B200Ppo* p, int M, cudaStream_t st) {
    float* ws = p->ws;
    const Workspace& w = p->w;
    int rc;
    if (g_chain) {
        if ((rc = chain_forward(p, critic_ptrs(p, M), actor_ptrs(p, 0), st))) return rc;
        k_value_head<<<(int)(((size_t)M * 4 + 255) / 256), 256, 0, st>>>(ws + w.C3, p->P(P_CW3), p->P(P_CB3), M, ws + w.V);
        g_launches += 1;
        return launch_status("k_value_head");
    }
    if ((rc = tc_fwd(p, ws + w.Xc, 61, 64, ws + w.Wc0h, ws + w.Wc0l, 64, p->P(P_CB0), ws + w.C1, M, 256, st))) return rc;
    if ((rc = tc_fwd(p, ws + w.C1, 256, 256, ws + w.Wc1h, ws + w.Wc1l, 256, p->P(P_CB1), ws + w.C2, M, 256, st))) return rc;
    if ((rc = tc_fwd(p, ws + w.C2, 256, 256, ws + w.Wc2h, ws + w.Wc2l, 256, p->P(P_CB2), ws + w.C3, M, 128, st))) return rc;
    k_value_head<<<(int)(((size_t)M * 4 + 255) / 256), 256, 0, st>>>(ws + w.C3, p->P(P_CW3), p->P(P_CB3), M, ws + w.V);
    g_launches += 1;
    return launch_status("k_value_head");
}

// =====================================================================================================================
extern "C" {

int b200_ppo_num_params(void) { return NPARAMS_PADDED; }
int b200_ppo_num_param_tensors(void) { return P_COUNT; }
int b200_ppo_param_info(int idx, const char** name, int* offset, int* rows, int* cols) {
    if (idx < 0 || idx >= P_COUNT) return set_error(B200_ERR_ARG, "b200_ppo_param_info: bad index");
    if (name) *name = kParams[idx].name;
    if (offset) *offset = kParams[idx].offset;
    if (rows) *rows = kParams[idx].rows;
    if (cols) *cols = kParams[idx].cols;
    return B200_OK;
}

int64_t b200_ppo_workspace_bytes(int horizon, int num_envs) {
    if (horizon <= 0 || num_envs <= 0) return 0;
    return (int64_t)(make_workspace(horizon, num_envs).total * sizeof(float));
}

int b200_ppo_create(const B200PpoConfig* cfg, float* params, float* grads, float* adam_m, float* adam_v, float* scalars,
                    double* dstats, void* workspace, int device, B200Ppo** out) {
    if (!cfg || !params || !grads || !adam_m || !adam_v || !scalars || !dstats || !workspace || !out)
        return set_error(B200_ERR_ARG, "b200_ppo_create: null pointer");
    if (cfg->horizon <= 0 || cfg->num_envs <= 0 || cfg->world_size <= 0) return set_error(B200_ERR_ARG, "b200_ppo_create: bad sizes");
    if ((reinterpret_cast<uintptr_t>(params) & 15) || (reinterpret_cast<uintptr_t>(grads) & 15) ||
        (reinterpret_cast<uintptr_t>(workspace) & 1023))
        return set_error(B200_ERR_ARG, "b200_ppo_create: params/grads must be 16-byte and workspace 1024-byte aligned");
    CUDA_TRY(cudaSetDevice(device));
    B200Ppo* p = new (std::nothrow) B200Ppo();
    if (!p) return set_error(B200_ERR_ARG, "out of host memory");
    p->cfg = *cfg;
    p->device = device;
    p->params = params; p->grads = grads; p->adam_m = adam_m; p->adam_v = adam_v; p->scalars = scalars;
    p->dstats = dstats;
    p->ws = (float*)workspace;
    p->w = make_workspace(cfg->horizon, cfg->num_envs);
    p->act_ctr = nullptr;
    p->maps = new (std::nothrow) tc::MapCache();
    if (!p->maps) { delete p; return set_error(B200_ERR_ARG, "out of host memory"); }
    p->num_sms = 148;
    cudaDeviceGetAttribute(&p->num_sms, cudaDevAttrMultiProcessorCount, device);
    p->peers = 0;
    p->xch_counter = nullptr;
    cudaError_t ce = cudaMalloc(&p->act_ctr, sizeof(unsigned long long));
    if (ce == cudaSuccess) ce = cudaMemset(p->act_ctr, 0, sizeof(unsigned long long));
    if (ce == cudaSuccess) ce = cudaMalloc(&p->xch_counter, 4 * sizeof(unsigned int));
    if (ce == cudaSuccess) ce = cudaMemset(p->xch_counter, 0, 4 * sizeof(unsigned int));
    if (ce != cudaSuccess) { delete p; return set_cuda_error(ce, "b200_ppo_create: counter allocation"); }
    *out = p;
    return B200_OK;
}
int b200_ppo_destroy(B200Ppo* p) {
    if (p && p->act_ctr) cudaFree(p->act_ctr);
    if (p && p->xch_counter) cudaFree(p->xch_counter);
    if (p) delete p->maps;
    delete p;
    return B200_OK;
}

#define NEED_PPO(p) if (!(p)) return set_error(B200_ERR_ARG, "null learner handle")

int b200_policy_act(B200Ppo* p, const float* obs, int n, float* actions, float* mu_out, const float* eps_in,
                    uint64_t seed, uint64_t step, int deterministic, void* stream) {
    NEED_PPO(p);
    if (!obs || !actions || n <= 0) return set_error(B200_ERR_ARG, "b200_policy_act: bad argument");
    cudaStream_t st = (cudaStream_t)stream;
    static unsigned long long configured = 0;   // per-device bit mask
    constexpr int SMEM = PF_SMEM_FLOATS * (int)sizeof(float);
    CUDA_TRY(ensure_dynamic_smem(k_policy_fused, SMEM, configured));
    const bool auto_step = (step == B200_STEP_AUTO);
    const bool reuse = (deterministic & B200_ACT_REUSE_WEIGHTS) != 0;
    PolicyFusedArgs a{};
    a.obs = obs;
    a.W0 = p->P(P_AW0); a.b0 = p->P(P_AB0); a.W1 = p->P(P_AW1); a.b1 = p->P(P_AB1); a.W2 = p->P(P_AW2); a.b2 = p->P(P_AB2);
    a.W3 = p->P(P_AW3); a.b3 = p->P(P_AB3); a.logstd = p->P(P_LOGSTD);
    a.eps_in = eps_in; a.ctr = auto_step ? p->act_ctr : nullptr; a.actions = actions; a.mu_out = mu_out;
    a.seed = seed; a.step = step; a.n = n; a.env_base = p->cfg.env_base; a.deterministic = deterministic & 1;
    if (g_h2_chain && n >= g_policy_tc_min_rows && n <= p->cfg.num_envs) {
        // tensor-core path: obs -> h2 words, hidden layers on the fused forward chain (the N-row buffers of b200_critic_value), head + sampling
        float* ws = p->ws;
        const Workspace& w = p->w;
        k_pack_inputs<<<(int)(((size_t)n * 64 + 255) / 256), 256, 0, st>>>(obs, nullptr, n, PackOut{nullptr, ws + w.LXh, nullptr, nullptr, nullptr, nullptr, 1});
        g_launches += 1;
        int rc;
        // the weight operands (B' / B'' words, pre-scaled biases) are rebuilt unless the caller vouches that the parameters have not
        // changed since its previous call (Runner.rollout: every step but the first) - a per-call argument, so a captured graph replays it
        if (!reuse && (rc = weight_prep(p, st)) != B200_OK) return rc;
        ChainNetPtrs c = actor_ptrs(p, n);
        c.Xh = ws + w.LXh; c.Xl = ws + w.LXl; c.H1 = ws + w.L1; c.H2 = ws + w.L2; c.H3 = ws + w.L3;
        if ((rc = chain_forward(p, critic_ptrs(p, 0), c, st)) != B200_OK) return rc;
        const int blocks = (n + 7) / 8;
        k_act_head<<<blocks < 4 * p->num_sms ? blocks : 4 * p->num_sms, 256, 0, st>>>(ws + w.L3, a);
        if (auto_step) k_bump<<<1, 1, 0, st>>>(p->act_ctr);
        g_launches += auto_step ? 2 : 1;
        return launch_status("k_act_head");
    }
    k_policy_fused<<<(n + PF_ROWS - 1) / PF_ROWS, PF_THREADS, SMEM, st>>>(a);
    if (auto_step) k_bump<<<1, 1, 0, st>>>(p->act_ctr);
    g_launches += auto_step ? 2 : 1;
    return launch_status("k_policy_fused");
}

int b200_critic_value(B200Ppo* p, const float* obs, const float* priv, int n, float* values, void* stream) {
    NEED_PPO(p);
    if (!obs || !priv || !values || n <= 0 || n > p->cfg.num_envs) return set_error(B200_ERR_ARG, "b200_critic_value: bad argument");
    cudaStream_t st = (cudaStream_t)stream;
    float* ws = p->ws;
    const Workspace& w = p->w;
    k_pack_inputs<<<(int)(((size_t)n * 64 + 255) / 256), 256, 0, st>>>(obs, priv, n, PackOut{nullptr, nullptr, nullptr, nullptr, ws + w.LXh, ws + w.LXl, g_h2_chain ? 1 : 0});
    g_launches += 1;
    int rc = weight_prep(p, st);   // the parameters may have changed since the last epoch
    if (rc != B200_OK) return rc;
    ChainNetPtrs c = critic_ptrs(p, n);
    c.Xh = ws + w.LXh; c.Xl = ws + w.LXl; c.H1 = ws + w.L1; c.H2 = ws + w.L2; c.H3 = ws + w.L3;
    if ((rc = chain_forward(p, c, actor_ptrs(p, 0), st)) != B200_OK) return rc;
    k_value_head<<<(int)(((size_t)n * 4 + 255) / 256), 256, 0, st>>>(ws + w.L3, p->P(P_CW3), p->P(P_CB3), n, values);
    g_launches += 1;
    return launch_status("k_value_head");
}

int b200_ppo_old_dist(B200Ppo* p, const float* obses, const float* privs, const float* actions, float* old_mu,
                      float* old_logp, void* stream) {
    NEED_PPO(p);
    if (!obses || !privs || !actions || !old_mu || !old_logp) return set_error(B200_ERR_ARG, "b200_ppo_old_dist: null pointer");
    cudaStream_t st = (cudaStream_t)stream;
    float* ws = p->ws;
    const int M = p->cfg.horizon * p->cfg.num_envs;
    p->last_staged = false;   // a new rollout: the next b200_ppo_epoch_a stages its last_obs / last_priv again
    k_pack_inputs<<<(int)(((size_t)M * 64 + 255) / 256), 256, 0, st>>>(
        obses, privs, M, PackOut{ws + p->w.Xa, ws + p->w.Xah, ws + p->w.Xal, ws + p->w.Xc, ws + p->w.Xch, ws + p->w.Xcl, g_h2_chain ? 1 : 0});
    int rc = weight_prep(p, st);
    if (rc != B200_OK) return rc;
    if ((rc = actor_forward_tc(p, M, st)) != B200_OK) return rc;
    k_old_logp<<<(M + 255) / 256, 256, 0, st>>>(ws + p->w.MU, actions, p->P(P_LOGSTD), M, old_mu, old_logp, p->scalars);
    g_launches += 2;  // + k_pack_inputs
    return launch_status("k_old_logp");
}

int b200_gae(float* rewards, const uint8_t* dones, const uint8_t* time_outs, const float* values, const float* last_values,
             double gamma, double lam, float* advantages, float* returns, double* stats, int horizon, int num_envs,
             void* stream) {
    if (!rewards || !dones || !time_outs || !values || !last_values || !advantages || !returns || horizon <= 0 || num_envs <= 0)
        return set_error(B200_ERR_ARG, "b200_gae: bad argument");
    k_gae<<<(num_envs + 31) / 32, 32, 0, (cudaStream_t)stream>>>(rewards, dones, time_outs, values, last_values,
                                                                      (float)gamma, (float)(gamma * lam), horizon,
                                                                      num_envs, advantages, returns, stats);
    g_launches += 1;
    return launch_status("k_gae");
}

int b200_ppo_epoch_a(B200Ppo* p, float* rewards, const uint8_t* dones, const uint8_t* time_outs, const float* last_obs,
                     const float* last_priv, void* stream) {
    NEED_PPO(p);
    if (!rewards || !dones || !time_outs || !last_obs || !last_priv) return set_error(B200_ERR_ARG, "b200_ppo_epoch_a: null pointer");
    cudaStream_t st = (cudaStream_t)stream;
    float* ws = p->ws;
    const int T = p->cfg.horizon, N = p->cfg.num_envs, M = T * N;
    int rc = weight_prep(p, st, p->dstats);  // the parameters changed in the previous epoch's b200_ppo_apply; also clears dstats
    if (rc != B200_OK) return rc;
    // last_values = critic(post-rollout obs): appended as rows [M, M+N) of the critic batch, evaluated in the same GEMMs.  The reference
    // evaluates est_value on the SAME tensors in every mini-epoch (utils/runner.py:132): they are staged by the first epoch after
    // b200_ppo_old_dist (or whenever the pointers change) and reused by the following ones
    if (!(p->last_staged && p->last_obs_ptr == last_obs && p->last_priv_ptr == last_priv && p->last_staged_h2 == g_h2_chain)) {
        const size_t o = (size_t)M * 64;
        k_pack_inputs<<<(int)(((size_t)N * 64 + 255) / 256), 256, 0, st>>>(
            last_obs, last_priv, N, PackOut{nullptr, nullptr, nullptr, ws + p->w.Xc + o, ws + p->w.Xch + o, ws + p->w.Xcl + o, g_h2_chain ? 1 : 0});
        g_launches += 1;
        p->last_staged = true; p->last_obs_ptr = last_obs; p->last_priv_ptr = last_priv; p->last_staged_h2 = g_h2_chain;
    }
    if (g_chain) {
        // both nets in ONE persistent launch: 800 critic + 768 actor tiles balance over the SMs better than two launches of ~5.3
        // waves each; the actor's mu is not needed before epoch_b's loss
        if ((rc = chain_forward(p, critic_ptrs(p, M + N), actor_ptrs(p, M), st)) != B200_OK) return rc;
        if ((rc = head_forward<1>(p, ws + p->w.C3, p->P(P_CW3), p->P(P_CB3), M + N, ws + p->w.V, st)) != B200_OK) return rc;
        if ((rc = head_forward<12>(p, ws + p->w.A3, p->P(P_AW3), p->P(P_AB3), M, ws + p->w.MU, st)) != B200_OK) return rc;
        p->actor_fwd_done = true;
    } else if ((rc = critic_forward_tc(p, M + N, st)) != B200_OK) return rc;
    k_gae<<<(N + 31) / 32, 32, 0, st>>>(rewards, dones, time_outs, ws + p->w.V, ws + p->w.V + M, (float)p->cfg.gamma,
                                           (float)(p->cfg.gamma * p->cfg.lam), T, N, ws + p->w.ADV, ws + p->w.RET, p->dstats);
    g_launches += 1;  // k_gae (the zeroing kernel is counted by weight_prep)
    if (p->peers) {   // this rank's advantage moments -> peers (summed in epoch_b, behind the actor forward)
        k_xchg_post_stats<<<1, 32, 0, st>>>(p->px, p->dstats, p->xch_counter);
        g_launches += 1;
    }
    return launch_status("k_gae");
}

int b200_ppo_epoch_b(B200Ppo* p, const float* actions, const float* old_mu, const float* old_logp, void* stream) {
    NEED_PPO(p);
    if (!actions || !old_mu || !old_logp) return set_error(B200_ERR_ARG, "b200_ppo_epoch_b: null pointer");
    if ((reinterpret_cast<uintptr_t>(actions) & 15) || (reinterpret_cast<uintptr_t>(old_mu) & 15))
        return set_error(B200_ERR_ARG, "b200_ppo_epoch_b: actions / old_mu must be 16-byte aligned (rows of 12 floats are read as float4)");
    cudaStream_t st = (cudaStream_t)stream;
    float* ws = p->ws;
    const Workspace& w = p->w;
    const int M = p->cfg.horizon * p->cfg.num_envs;
    float *MU = ws + w.MU, *DV = ws + w.DV, *DMU = ws + w.DMU;
    float *G1 = ws + w.G1, *G2 = ws + w.G2;
    WgradJobs jobs{};
    int rc = B200_OK;
    if (!p->actor_fwd_done && (rc = actor_forward_tc(p, M, st)) != B200_OK) return rc;
    p->actor_fwd_done = false;
    CUDA_TRY(cudaMemsetAsync(p->grads, 0, NPARAMS_PADDED * sizeof(float), st));
    if (p->peers) {   // global advantage normalisation (utils/runner.py:145 over all ranks' samples)
        k_xchg_reduce_stats<<<1, 32, 0, st>>>(p->px, p->dstats, p->xch_counter);
        g_launches += 1;
    }
    k_loss<<<(M + LOSS_BLOCK - 1) / LOSS_BLOCK, LOSS_BLOCK, 0, st>>>(ws + w.V, ws + w.RET, ws + w.ADV, MU, actions, old_mu, old_logp,
                                                                    p->P(P_LOGSTD), p->scalars, p->dstats, M, p->cfg.e_clip,
                                                                    p->cfg.bound_coef, DV, DMU, reinterpret_cast<unsigned int*>(ws + w.ZR));
    k_finalize_loss<<<1, 256, 0, st>>>(p->dstats, p->cfg.entropy_coef, p->G(P_LOGSTD), p->P(P_AW3), p->P(P_CW3),
                                       reinterpret_cast<const unsigned int*>(ws + w.ZR), ws + w.ZR + 64, ws + w.SC);
    g_launches += 3;  // memset, k_loss, k_finalize_logstd
    if ((rc = launch_status("k_loss")) != B200_OK) return rc;
    if (g_chain) {
        // head backward kernels produce dz3 of both nets (+ the head's own gradients and layer 3's bias gradient), ONE fused chain
        // launch produces dz2, dz1 and the remaining bias gradients, then the six weight-gradient GEMMs
        float *GA3 = ws + w.GA3, *GA2 = ws + w.GA2, *GA1 = ws + w.GA1, *GC3 = ws + w.GC3, *GC2 = ws + w.GC2, *GC1 = ws + w.GC1;
        if (g_h2_chain) {
            // streaming head backward (heads.cuh): bulk-copy fed, dz3 written as h2 words
            static unsigned long long cfg_a = 0, cfg_c = 0;
            CU_TRY(ensure_dynamic_smem(heads::k_head_bwd_pipe<12>, heads::HeadSmem<12>::TOTAL, cfg_a));
            CU_TRY(ensure_dynamic_smem(heads::k_head_bwd_pipe<1>, heads::HeadSmem<1>::TOTAL, cfg_c));
            const int tiles = (M + heads::HB2_ROWS - 1) / heads::HB2_ROWS, grid = tiles < 2 * p->num_sms ? tiles : 2 * p->num_sms;
            heads::k_head_bwd_pipe<12><<<grid, heads::HB2_THREADS, heads::HeadSmem<12>::TOTAL, st>>>(
                ws + w.A3, p->P(P_AW3), DMU, M, reinterpret_cast<uint32_t*>(GA3), p->G(P_AW3), p->G(P_AB3), p->G(P_AB2), ws + w.SC + h2::SC_SG_A);
            heads::k_head_bwd_pipe<1><<<grid, heads::HB2_THREADS, heads::HeadSmem<1>::TOTAL, st>>>(
                ws + w.C3, p->P(P_CW3), DV, M, reinterpret_cast<uint32_t*>(GC3), p->G(P_CW3), p->G(P_CB3), p->G(P_CB2), ws + w.SC + h2::SC_SG_C);
        } else {
            k_actor_head_bwd<<<(M + HB_ROWS - 1) / HB_ROWS, HB_THREADS, 0, st>>>(ws + w.A3, p->P(P_AW3), DMU, M, GA3, p->G(P_AW3), p->G(P_AB3), p->G(P_AB2), nullptr);
            k_value_head_bwd<<<(M + HB_ROWS - 1) / HB_ROWS, HB_THREADS, 0, st>>>(ws + w.C3, p->P(P_CW3), DV, M, GC3, p->G(P_CW3), p->G(P_CB3), p->G(P_CB2), nullptr);
        }
        g_launches += 2;
        if ((rc = launch_status("k_head_bwd")) != B200_OK) return rc;
        if ((rc = chain_backward(p, M, true, true, st))) return rc;
        if (g_h2_chain) {
            // the chains wrote every operand as h2 words: the weight-gradient kernel takes them as they are
            const WgH2Operand ops[6] = {{GA3, ws + w.A2, 128, 128, 128, P_AW2, true, h2::S_ACT},  {GA2, ws + w.A1, 128, 256, 256, P_AW1, true, h2::S_ACT},
                                        {GA1, ws + w.Xah, 256, 64, 47, P_AW0, true, h2::S_X},      {GC3, ws + w.C2, 128, 256, 256, P_CW2, false, h2::S_ACT},
                                        {GC2, ws + w.C1, 256, 256, 256, P_CW1, false, h2::S_ACT},  {GC1, ws + w.Xch, 256, 64, 61, P_CW0, false, h2::S_X}};
            return wgrad_h2(p, ops, M, st);
        }
        if ((rc = tc_wgrad(p, jobs, 0, GA3, 128, ws + w.A2, 128, 128, 128, p->G(P_AW2), M, st))) return rc;
        if ((rc = tc_wgrad(p, jobs, 1, GA2, 128, ws + w.A1, 256, 256, 256, p->G(P_AW1), M, st))) return rc;
        if ((rc = tc_wgrad(p, jobs, 2, GA1, 256, ws + w.Xa, 64, 64, 47, p->G(P_AW0), M, st))) return rc;
        if ((rc = tc_wgrad(p, jobs, 3, GC3, 128, ws + w.C2, 256, 256, 256, p->G(P_CW2), M, st))) return rc;
        if ((rc = tc_wgrad(p, jobs, 4, GC2, 256, ws + w.C1, 256, 256, 256, p->G(P_CW1), M, st))) return rc;
        if ((rc = tc_wgrad(p, jobs, 5, GC1, 256, ws + w.Xc, 64, 64, 61, p->G(P_CW0), M, st))) return rc;
        return wgrad_reduce(jobs, st);
    }
    // ---- actor backward: the 12-wide head as a fused FMA kernel, the hidden layers on tcgen05
    k_actor_head_bwd<<<(M + HB_ROWS - 1) / HB_ROWS, HB_THREADS, 0, st>>>(ws + w.A3, p->P(P_AW3), DMU, M, G1, p->G(P_AW3), p->G(P_AB3), p->G(P_AB2), nullptr);
    g_launches += 1;
    if ((rc = launch_status("k_actor_head_bwd")) != B200_OK) return rc;
    // (each bias gradient = column sum of the layer's output gradient, accumulated by the kernel that produces it)
    if ((rc = tc_wgrad(p, jobs, 0, G1, 128, ws + w.A2, 128, 128, 128, p->G(P_AW2), M, st))) return rc;
    if ((rc = tc_dgrad(p, G1, 128, ws + w.Wa2Th, ws + w.Wa2Tl, 128, ws + w.A2, G2, p->G(P_AB1), M, st))) return rc;
    if ((rc = tc_wgrad(p, jobs, 1, G2, 128, ws + w.A1, 256, 256, 256, p->G(P_AW1), M, st))) return rc;
    if ((rc = tc_dgrad(p, G2, 128, ws + w.Wa1Th, ws + w.Wa1Tl, 256, ws + w.A1, G1, p->G(P_AB0), M, st))) return rc;
    if ((rc = tc_wgrad(p, jobs, 2, G1, 256, ws + w.Xa, 64, 64, 47, p->G(P_AW0), M, st))) return rc;
    // ---- critic backward
    k_value_head_bwd<<<(M + HB_ROWS - 1) / HB_ROWS, HB_THREADS, 0, st>>>(ws + w.C3, p->P(P_CW3), DV, M, G1, p->G(P_CW3), p->G(P_CB3), p->G(P_CB2), nullptr);
    g_launches += 1;
    if ((rc = launch_status("k_value_head_bwd")) != B200_OK) return rc;
    if ((rc = tc_wgrad(p, jobs, 3, G1, 128, ws + w.C2, 256, 256, 256, p->G(P_CW2), M, st))) return rc;
    if ((rc = tc_dgrad(p, G1, 128, ws + w.Wc2Th, ws + w.Wc2Tl, 256, ws + w.C2, G2, p->G(P_CB1), M, st))) return rc;
    if ((rc = tc_wgrad(p, jobs, 4, G2, 256, ws + w.C1, 256, 256, 256, p->G(P_CW1), M, st))) return rc;
    if ((rc = tc_dgrad(p, G2, 256, ws + w.Wc1Th, ws + w.Wc1Tl, 256, ws + w.C1, G1, p->G(P_CB0), M, st))) return rc;
    if ((rc = tc_wgrad(p, jobs, 5, G1, 256, ws + w.Xc, 64, 64, 61, p->G(P_CW0), M, st))) return rc;
    return wgrad_reduce(jobs, st);
}

int b200_ppo_epoch(B200Ppo* p, const float* actions, float* rewards, const uint8_t* dones, const uint8_t* time_outs,
                   const float* last_obs, const float* last_priv, const float* old_mu, const float* old_logp, void* stream) {
    const int rc = b200_ppo_epoch_a(p, rewards, dones, time_outs, last_obs, last_priv, stream);
    if (rc != B200_OK) return rc;
    return b200_ppo_epoch_b(p, actions, old_mu, old_logp, stream);
}

int b200_ppo_apply(B200Ppo* p, void* stream) {
    NEED_PPO(p);
    cudaStream_t st = (cudaStream_t)stream;
    const float inv_world = 1.0f / (float)p->cfg.world_size;
    if (p->peers) {   // sum-all-reduce of the gradient + loss sums over NVLink peer memory, fused with the gradient norm
        k_xchg_post_grads<<<64, 256, 0, st>>>(p->px, p->grads, p->dstats, p->xch_counter);
        k_xchg_reduce_grads<<<96, 256, 0, st>>>(p->px, p->grads, p->dstats, p->xch_counter, inv_world);
        g_launches += 1;
    } else {
        k_grad_sumsq<<<148, 256, 0, st>>>(p->grads, NPARAMS_PADDED, inv_world, p->dstats);
    }
    k_adam<<<(NPARAMS_PADDED + 255) / 256, 256, 0, st>>>(p->params, p->grads, p->adam_m, p->adam_v, NPARAMS_PADDED, p->scalars,
                                                         p->dstats, inv_world, p->cfg.max_grad_norm, p->cfg.adam_beta1,
                                                         p->cfg.adam_beta2, p->cfg.adam_eps, p->xch_counter + 3, p->cfg.desired_kl,
                                                         p->cfg.lr_min, p->cfg.lr_max, p->cfg.lr_factor);
    g_launches += 2;
    return launch_status("b200_ppo_apply");
}

long long b200_launch_count(void) { return g_launches; }

long long b200_ppo_peer_buffer_bytes(void) { return (long long)(XCH_FLAG_FLOATS + 2 * (size_t)XCH_SLOT_FLOATS) * (long long)sizeof(float); }
int b200_ppo_bind_peers(B200Ppo* p, const unsigned long long* buffer_ptrs, int rank, int world) {
    NEED_PPO(p);
    if (!buffer_ptrs || world < 1 || world > XCH_MAX_WORLD || rank < 0 || rank >= world || world != p->cfg.world_size)
        return set_error(B200_ERR_ARG, "b200_ppo_bind_peers: bad argument (world must equal the learner's world_size, <= 16)");
    for (int r = 0; r < world; ++r) {
        if (!buffer_ptrs[r] || (buffer_ptrs[r] & 15ull)) return set_error(B200_ERR_ARG, "b200_ppo_bind_peers: null / misaligned peer buffer");
        p->px.bufs[r] = buffer_ptrs[r];
    }
    p->px.rank = rank;
    p->px.world = world;
    p->peers = 1;
    CUDA_TRY(cudaMemset(p->xch_counter, 0, 4 * sizeof(unsigned int)));   // the flags of a fresh peer buffer are zero: sequence numbers restart
    return B200_OK;
}
int b200_tc_set_pair(int enable) {
    g_tc_pair = enable != 0;
    return B200_OK;
}
int b200_tc_set_h2(int mode) {
    g_h2_chain = (mode & 1) != 0;
    if (g_h2_chain) { g_chain = true; g_chain_pair = false; }
    if (mode & 0x30) g_h2_groups_fwd = ((mode >> 4) & 3) >= 2 ? 2 : 1;   // bits 4-5 / 6-7: epilogue warp groups (1 or 2) of the forward /
    if (mode & 0xC0) g_h2_groups_bwd = ((mode >> 6) & 3) >= 2 ? 2 : 1;   // backward chain
    if (mode & 0x300) g_h2_mc = ((mode >> 8) & 3) >= 2;                  // bits 8-9: 1 = every CTA streams its own weights, 2 = multicast in CTA pairs
    if (mode & 0xC00) {   // bits 10-11: b200_policy_act 1 = always FP32-FMA, 2 = always the tensor-core chain, 3 = by row count (default)
        const int m = (mode >> 10) & 3;
        g_policy_tc_min_rows = m == 1 ? 0x7fffffff : (m == 2 ? 1 : (getenv("B200_POLICY_TC_MIN_ROWS") ? atoi(getenv("B200_POLICY_TC_MIN_ROWS")) : 2048));
    }
    return B200_OK;
}
int b200_tc_set_chain(int enable) {
    g_chain = enable != 0;
    g_chain_pair = enable == 2;
    if (enable != 1) g_h2_chain = false;   // the h2 kernels exist as single-CTA fused chains only (re-enable with b200_tc_set_h2(1))
    return B200_OK;
}

#ifdef B200_TC_TIMELINE
/* debug builds only (tools/tc_timeline.py): per-CTA timelines of the tcgen05 GEMM launches since the last reset, in launch order */
int b200_tc_timeline_reset(void) { g_tl_slot = 0; return B200_OK; }
int b200_tc_timeline_read(unsigned long long* out /* [TL_SLOTS][160][12] */) {
    CUDA_TRY(cudaDeviceSynchronize());
    CUDA_TRY(cudaMemcpyFromSymbol(out, tc::g_tl, sizeof(unsigned long long) * TL_SLOTS * 160 * 12));
    return g_tl_slot;
}
#endif

#ifdef B200_CHAIN_TL
/* debug builds only (tools/chain_timeline.py): per-CTA blocked-cycle counters of the last k_mlp_fwd / k_mlp_bwd launch */
int b200_chain_timeline_read(long long* out /* [2][160][16] */) {
    CUDA_TRY(cudaDeviceSynchronize());
    CUDA_TRY(cudaMemcpyFromSymbol(out, chain::g_chain_tl, sizeof(long long) * 2 * 160 * 16));
    return B200_OK;
}
#endif

int b200_profile_gemm(int enable) {
    if (enable && !g_prof.ev) {
        g_prof.ev = new (std::nothrow) cudaEvent_t[2 * GemmProfile::kMax];
        g_prof.kind = new (std::nothrow) unsigned char[GemmProfile::kMax];
        if (!g_prof.ev || !g_prof.kind) return set_error(B200_ERR_ARG, "out of host memory");
        for (int i = 0; i < 2 * GemmProfile::kMax; ++i) CUDA_TRY(cudaEventCreate(&g_prof.ev[i]));
    }
    g_prof.on = enable != 0;
    g_prof.used = 0;
    for (int k = 0; k < PK_COUNT; ++k) g_prof.flops[k] = g_prof.bytes[k] = 0.0;
    return B200_OK;
}
/* kind: 0 = k_gemm3x (mma.sync), 1 = k_tc_rowmajor (tcgen05 fwd / dgrad), 2 = k_tc_wgrad (tcgen05), -1 = all */
int b200_profile_gemm_read(int kind, double* total_ms, double* total_flops, int* launches) {
    double ms = 0.0, fl = 0.0;
    int cnt = 0;
    for (int i = 0; i < g_prof.used; ++i) {
        if (kind >= 0 && g_prof.kind[i] != kind) continue;
        CUDA_TRY(cudaEventSynchronize(g_prof.ev[2 * i + 1]));
        float t = 0.f;
        CUDA_TRY(cudaEventElapsedTime(&t, g_prof.ev[2 * i], g_prof.ev[2 * i + 1]));
        ms += (double)t;
        cnt += 1;
    }
    for (int k = 0; k < PK_COUNT; ++k)
        if (kind < 0 || kind == k) fl += g_prof.flops[k];
    if (total_ms) *total_ms = ms;
    if (total_flops) *total_flops = fl;
    if (launches) *launches = cnt;
    return B200_OK;
}

int b200_profile_gemm_bytes(int kind, double* total_bytes) {
    double b = 0.0;
    for (int k = 0; k < PK_COUNT; ++k)
        if (kind < 0 || kind == k) b += g_prof.bytes[k];
    if (total_bytes) *total_bytes = b;
    return B200_OK;
}

float* b200_ppo_buffer(B200Ppo* p, int which) {
    if (!p) return nullptr;
    switch (which) {
        case 0: return p->ws + p->w.V;
        case 1: return p->ws + p->w.ADV;
        case 2: return p->ws + p->w.RET;
        case 3: return p->ws + p->w.MU;
        case 4: return p->ws + p->w.V + (size_t)p->cfg.horizon * p->cfg.num_envs;
        case 5: return p->ws + p->w.DV;
        case 6: return p->ws + p->w.DMU;
        case 7: return p->ws + p->w.C1;
        case 8: return p->ws + p->w.C2;
        case 9: return p->ws + p->w.C3;
        case 10: return p->ws + p->w.A1;
        case 11: return p->ws + p->w.A2;
        case 12: return p->ws + p->w.A3;
        case 13: return p->ws + p->w.GC2;
        case 14: return p->ws + p->w.GC1;
        case 15: return p->ws + p->w.GA2;
        case 16: return p->ws + p->w.GA1;
        case 17: return p->ws + p->w.GC3;
        case 18: return p->ws + p->w.GA3;
    }
    return nullptr;
}

}  // extern "C"
