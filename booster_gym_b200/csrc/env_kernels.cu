// env_kernels.cu - CUDA kernels + C-ABI for the T1 environment step (include/b200_t1.h, "env" half) except the physics
// loop (physics_kernels.cu): K2 post-physics (derived state, feet, kicks/pushes, termination, 23 reward terms, reset,
// command resampling, observations) and K3 reset/DR, one thread per environment over structure-of-arrays rows.
// Compiled with --fmad=false: every fp32 operation rounds separately, like the reference's eager torch ops, which is
// what makes the masks (feet_contact, reset, dof_pos_limits counts) bit-exact.  These passes are HBM-bound.
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <new>

#include "env_handle.cuh"

using namespace b200;
namespace b200 { extern long long g_launches; }

__global__ void k_advance(long long* ctr, long long common_step, int bump_common) {
    if (bump_common) ctr[1] = (common_step >= 0) ? common_step : ctr[1] + 1;
    const long long s = ctr[0] + 1;
    ctr[0] = s;
    ctr[2 + ((s + 1) & 1)] = 0;
}

#define POST_BLOCK 64
#define POST_ROW (B200_NOBS + B200_NPRIV)

// coalesced write-out of the [N,47] / [N,14] row-major API tensors through shared memory
__device__ __forceinline__ void store_obs_rows(float* sbuf, const float* obs_l, const float* priv_l, int e0, int n,
                                               float* __restrict__ obs, float* __restrict__ priv) {
    const int t = threadIdx.x;
#pragma unroll
    for (int i = 0; i < B200_NOBS; ++i) sbuf[t * POST_ROW + i] = obs_l[i];
#pragma unroll
    for (int i = 0; i < B200_NPRIV; ++i) sbuf[t * POST_ROW + B200_NOBS + i] = priv_l[i];
    __syncthreads();
    const int cnt = min(POST_BLOCK, n - e0);
    for (int idx = t; idx < cnt * B200_NOBS; idx += POST_BLOCK) {
        const int el = idx / B200_NOBS, i = idx - el * B200_NOBS;
        obs[(size_t)e0 * B200_NOBS + idx] = sbuf[el * POST_ROW + i];
    }
    for (int idx = t; idx < cnt * B200_NPRIV; idx += POST_BLOCK) {
        const int el = idx / B200_NPRIV, i = idx - el * B200_NPRIV;
        priv[(size_t)e0 * B200_NPRIV + idx] = sbuf[el * POST_ROW + B200_NOBS + i];
    }
}

// PHASE 0: the whole post-physics step.  PHASE 1 / 2: the two halves around the command-curriculum grid update (t1_env.cuh).
template <int PHASE>
__global__ void __launch_bounds__(POST_BLOCK)
k_post(EnvView v, const __grid_constant__ B200T1ModelF m, const __grid_constant__ B200T1Config c, TerrainView terr,
       const long long* __restrict__ ctr, int noise_on, float* __restrict__ obs, float* __restrict__ priv,
       float* __restrict__ rew, uint8_t* __restrict__ done, float* __restrict__ rew_terms, long long* flags,
       double* stats) {
    __shared__ float sbuf[POST_BLOCK * POST_ROW];
    const int e0 = blockIdx.x * POST_BLOCK;
    const int e = e0 + threadIdx.x;
    const long long step = ctr[0], common_step = ctr[1];
    float obs_l[B200_NOBS], priv_l[B200_NPRIV];
    if (e < v.n) {
        const StepOut o = env_post_physics<B200T1ModelF, PHASE>(v, e, m, c, terr, common_step, (uint64_t)step, noise_on, obs_l, priv_l, rew_terms);
        if (PHASE != 2) {
        rew[e] = o.rew;
        done[e] = (uint8_t)o.done;
        }
        if (PHASE != 2 && o.done) {
            flags[step & 1] = 1;  // benign race: every writer stores 1
            // device episode statistics (utils/recorder.py:36-62): flush this episode's sums
            float* f = v.f;
            int32_t* is = v.is;
            const int n = v.n;
            for (int k = 0; k < 1 + c.n_rew; ++k) {
                atomicAdd(&stats[k], (double)FS(F_episode_sums + k));
                FS(F_episode_sums + k) = 0.0f;
            }
            atomicAdd(&stats[1 + c.n_rew], (double)IS(I_episode_steps));
            atomicAdd(&stats[2 + c.n_rew], 1.0);
            IS(I_episode_steps) = 0;
        }
    }
    if (PHASE != 1) store_obs_rows(sbuf, obs_l, priv_l, e0, v.n, obs, priv);
}

// The whole post-physics step with one env on TWO warps (the default without the command curriculum).  k_post<0> runs ~12 000
// dependent instructions per env on one warp per SM: it is bound by instruction latency and fetch (ncu: 8.5 stall cycles per issue,
// nothing to overlap them with), not by memory.  The reward terms (R) and the reset + teleport + command resampling + observations
// (T) of an env are independent once part A (derived state, feet, termination) is done and the reward snapshot sits in registers
// (t1_env.cuh), so warp 0 runs A, snapshot, R and warp 1 runs [reset], T of the same 32 envs at the same time.  Every quantity is
// still computed by exactly one thread with the same instructions: results are bit-identical to k_post<0> (tests/test_gpu_env_post.py).
#define PAIR_ENVS 32
__global__ void __launch_bounds__(2 * PAIR_ENVS)
k_post_pair(EnvView v, const __grid_constant__ B200T1ModelF m, const __grid_constant__ B200T1Config c, TerrainView terr,
            const long long* __restrict__ ctr, int noise_on, float* __restrict__ obs, float* __restrict__ priv,
            float* __restrict__ rew, uint8_t* __restrict__ done, float* __restrict__ rew_terms, long long* flags, double* stats) {
    __shared__ float sbuf[PAIR_ENVS * POST_ROW];
    __shared__ int s_reset[PAIR_ENVS];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int e0 = blockIdx.x * PAIR_ENVS;
    const int e = e0 + lane;
    const long long step = ctr[0], common_step = ctr[1];
    const bool live = e < v.n;
    PostA a;
    a.h_base = 0.0f; a.finite = true; a.reset = false; a.time_out = false;
    RewardSnap snap;
    if (warp == 0 && live) {
        a = env_post_a<B200T1ModelF, 1>(v, e, m, c, terr, common_step, (uint64_t)step);   // (left foot; the right one on warp 1)
        s_reset[lane] = a.reset ? 1 : 0;
    } else if (live) {
        env_refresh_feet<B200T1ModelF, 2>(v, e, m, terr);   // reads the physics step's foot pose only: independent of part A
    }
    __syncthreads();   // part A's state (global) and the reset flags are visible to both warps
    if (warp == 0 && live) reward_snapshot(v, e, snap);
    __syncthreads();   // ... and the snapshot is taken before warp 1 starts overwriting what it reads
    if (warp == 0) {
        if (live) {
            const float r = env_post_rewards(v, e, c, snap, a.h_base, a.finite, rew_terms);
            rew[e] = r;
            done[e] = (uint8_t)(a.reset ? 1 : 0);
            if (a.reset) {
                flags[step & 1] = 1;  // benign race: every writer stores 1
                // device episode statistics (utils/recorder.py:36-62): flush this episode's sums
                float* f = v.f;
                int32_t* is = v.is;
                const int n = v.n;
                for (int k = 0; k < 1 + c.n_rew; ++k) {
                    atomicAdd(&stats[k], (double)FS(F_episode_sums + k));
                    FS(F_episode_sums + k) = 0.0f;
                }
                atomicAdd(&stats[1 + c.n_rew], (double)IS(I_episode_steps));
                atomicAdd(&stats[2 + c.n_rew], 1.0);
                IS(I_episode_steps) = 0;
            }
        }
    } else if (live) {
        float obs_l[B200_NOBS], priv_l[B200_NPRIV];
        if (s_reset[lane]) env_reset_one(v, e, c, terr, (uint64_t)step);
        env_post_tail(v, e, m, c, terr, (uint64_t)step, noise_on, obs_l, priv_l);
#pragma unroll
        for (int i = 0; i < B200_NOBS; ++i) sbuf[lane * POST_ROW + i] = obs_l[i];
#pragma unroll
        for (int i = 0; i < B200_NPRIV; ++i) sbuf[lane * POST_ROW + B200_NOBS + i] = priv_l[i];
    }
    __syncthreads();
    // coalesced write-out of the [N,47] / [N,14] row-major API tensors
    const int cnt = min(PAIR_ENVS, v.n - e0), t = threadIdx.x;
    for (int idx = t; idx < cnt * B200_NOBS; idx += 2 * PAIR_ENVS) {
        const int el = idx / B200_NOBS, i = idx - el * B200_NOBS;
        obs[(size_t)e0 * B200_NOBS + idx] = sbuf[el * POST_ROW + i];
    }
    for (int idx = t; idx < cnt * B200_NPRIV; idx += 2 * PAIR_ENVS) {
        const int el = idx / B200_NPRIV, i = idx - el * B200_NPRIV;
        priv[(size_t)e0 * B200_NPRIV + idx] = sbuf[el * POST_ROW + B200_NOBS + i];
    }
}

// envs/t1.py:404-413 for the whole grid + the running sums the draws of :416 use (one block, one thread per cell)
__global__ void k_curriculum_apply(float* __restrict__ prob, int32_t* __restrict__ count, float* __restrict__ cdf, int cells, float rate) {
    for (int i = threadIdx.x; i < cells; i += blockDim.x) curriculum_apply_cell(prob, count, i, rate);
    __syncthreads();
    if (threadIdx.x == 0) curriculum_scan(prob, cdf, cells);
}

// extras["time_outs"] is rebound only on steps where at least one env resets (envs/t1.py:317 vs :556, SURVEY 8a note 1)
__global__ void k_finalize_timeouts(const int32_t* __restrict__ is, int n, const long long* __restrict__ ctr,
                                    const long long* __restrict__ flags, uint8_t* __restrict__ time_out) {
    const int e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= n) return;
    if (flags[ctr[0] & 1]) time_out[e] = (uint8_t)is[(size_t)I_time_out_buf * n + e];
}

template <int PHASE>   // 0 = all; 1 = the resets, 2 = commands + observations (around the command-curriculum grid update)
__global__ void __launch_bounds__(POST_BLOCK)
k_reset_all(EnvView v, const __grid_constant__ B200T1ModelF m, const __grid_constant__ B200T1Config c, TerrainView terr,
            const long long* __restrict__ ctr, float* __restrict__ obs, float* __restrict__ priv) {
    __shared__ float sbuf[POST_BLOCK * POST_ROW];
    const int e0 = blockIdx.x * POST_BLOCK;
    const int e = e0 + threadIdx.x;
    const uint64_t step = (uint64_t)ctr[0];
    float obs_l[B200_NOBS], priv_l[B200_NPRIV];
    if (e < v.n) {
        // T1.reset(): _reset_idx(all) ; _resample_commands ; _compute_observations   (envs/t1.py:294-299)
        if (PHASE != 2) env_reset_one(v, e, c, terr, step);
        float* f = v.f;
        int32_t* is = v.is;
        const int n = v.n;
        if (PHASE != 1) {
            if (IS(I_episode_length_buf) == IS(I_cmd_resample_time)) env_resample_command(v, e, c, step);
            env_observations(v, e, c, terr, step, 1, obs_l, priv_l);
        }
    }
    if (PHASE != 1) store_obs_rows(sbuf, obs_l, priv_l, e0, v.n, obs, priv);
}

__global__ void k_init_params(EnvView v, const __grid_constant__ B200T1ModelF m, const __grid_constant__ B200T1Config c) {
    const int e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e < v.n) env_init_params(v, e, m, c);
}

__global__ void k_terrain_heights(TerrainView terr, const float* __restrict__ xy, int stride, int count, float* __restrict__ out) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < count) out[i] = terr(xy[(size_t)i * stride], xy[(size_t)i * stride + 1]);
}

__global__ void k_rng_fill(uint64_t seed, int env_base, int n, uint64_t step, int purpose, int sub, int kind, float* out) {
    const int e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= n) return;
    const Philox4 p = rng_words(seed, (uint32_t)(env_base + e), step, purpose, sub);
    const Rand4 r = rand4(p);
#pragma unroll
    for (int l = 0; l < 4; ++l) {
        float val = (kind == 0) ? r.u[l] : ((kind == 1) ? r.n[l] : __uint_as_float(p.w[l]));
        out[(size_t)l * n + e] = val;
    }
}

// ---------------------------------------------------------------------------------------------------------------
extern "C" {

int b200_t1_num_float_rows(void) { return F_ROWS; }
int b200_t1_num_int_rows(void) { return I_ROWS; }
int b200_t1_num_fields(int kind) {
    return kind == 0 ? (int)(sizeof(kFloatFields) / sizeof(FieldInfo)) : (int)(sizeof(kIntFields) / sizeof(FieldInfo));
}
int b200_t1_field_info(int kind, int idx, const char** name, int* row, int* count) {
    if (idx < 0 || idx >= b200_t1_num_fields(kind)) return B200_ERR_ARG;
    const FieldInfo& fi = (kind == 0) ? kFloatFields[idx] : kIntFields[idx];
    if (name) *name = fi.name;
    if (row) *row = fi.row;
    if (count) *count = fi.count;
    return B200_OK;
}

int b200_t1_create(const B200T1ModelF* model, const B200T1Config* cfg, const int16_t* hf_host, int hf_rows, int hf_cols,
                   int num_envs, int device, uint64_t seed, B200T1Handle** out) {
    if (!model || !cfg || !out || num_envs <= 0) return set_error(B200_ERR_ARG, "b200_t1_create: bad argument");
    if (cfg->terrain_type != 0 && cfg->terrain_type != 1) return set_error(B200_ERR_ARG, "Invalid terrain type");
    if (cfg->terrain_type == 1 && (!hf_host || hf_rows <= 1 || hf_cols <= 1))
        return set_error(B200_ERR_ARG, "trimesh terrain needs a heightfield");
    if (cfg->n_rew < 0 || cfg->n_rew > B200_MAX_REW) return set_error(B200_ERR_ARG, "bad reward count");
    if (cfg->decimation <= 0 || cfg->resample_hi <= cfg->resample_lo) return set_error(B200_ERR_ARG, "bad control/command config");
    if (cfg->curriculum && (cfg->cur_lin_levels < 0 || cfg->cur_ang_levels < 0 || (2 * cfg->cur_lin_levels + 1) * (2 * cfg->cur_ang_levels + 1) > 4096))
        return set_error(B200_ERR_ARG, "bad command-curriculum grid");
    CUDA_TRY(cudaSetDevice(device));
    B200T1Handle* h = new (std::nothrow) B200T1Handle();
    if (!h) return set_error(B200_ERR_ARG, "out of host memory");
    memset(h, 0, sizeof(*h));
    h->model = *model;
    h->cfg = *cfg;
    h->num_envs = num_envs;
    h->device = device;
    h->seed = seed;
    h->env_base = 0;
    h->total_envs = num_envs;
    h->hf_rows = hf_rows;
    h->hf_cols = hf_cols;
    if (cfg->terrain_type == 1) {
        const size_t bytes = (size_t)hf_rows * hf_cols * sizeof(int16_t);
        CUDA_TRY(cudaMalloc(&h->hf_dev, bytes));
        CUDA_TRY(cudaMemcpy(h->hf_dev, hf_host, bytes, cudaMemcpyHostToDevice));
        int16_t top = hf_host[0];
        for (size_t i = 1; i < (size_t)hf_rows * hf_cols; ++i) top = hf_host[i] > top ? hf_host[i] : top;
        h->hf_max = (float)((double)top * cfg->vertical_scale) + 1e-5f;  // bound of the bilinear interpolant (terrain.cuh)
    }
    CUDA_TRY(cudaMalloc(&h->ctr_dev, 4 * sizeof(long long)));
    CUDA_TRY(cudaMemset(h->ctr_dev, 0, 4 * sizeof(long long)));
    CUDA_TRY(cudaMalloc(&h->stats_dev, (B200_MAX_REW + 4) * sizeof(double)));
    CUDA_TRY(cudaMemset(h->stats_dev, 0, (B200_MAX_REW + 4) * sizeof(double)));
    if (cfg->curriculum) {
        h->cur_cells = (2 * cfg->cur_lin_levels + 1) * (2 * cfg->cur_ang_levels + 1);
        CUDA_TRY(cudaMalloc(&h->cur_count, h->cur_cells * sizeof(int32_t)));
        CUDA_TRY(cudaMemset(h->cur_count, 0, h->cur_cells * sizeof(int32_t)));
        CUDA_TRY(cudaMalloc(&h->cur_cdf, h->cur_cells * sizeof(float)));
    }
    *out = h;
    return B200_OK;
}

int b200_t1_destroy(B200T1Handle* h) {
    if (!h) return B200_OK;
    cudaSetDevice(h->device);
    if (h->hf_dev) cudaFree(h->hf_dev);
    if (h->ctr_dev) cudaFree(h->ctr_dev);
    if (h->stats_dev) cudaFree(h->stats_dev);
    if (h->cur_count) cudaFree(h->cur_count);
    if (h->cur_cdf) cudaFree(h->cur_cdf);
    delete h;
    return B200_OK;
}

int b200_t1_bind_state(B200T1Handle* h, float* fstate, int32_t* istate) {
    if (!h || !fstate || !istate) return set_error(B200_ERR_ARG, "b200_t1_bind_state: null pointer");
    h->fstate = fstate;
    h->istate = istate;
    return B200_OK;
}
int b200_t1_bind_curriculum(B200T1Handle* h, float* prob, int rows, int cols) {
    if (!h || !prob) return set_error(B200_ERR_ARG, "b200_t1_bind_curriculum: null pointer");
    if (!h->cfg.curriculum) return set_error(B200_ERR_STATE, "b200_t1_bind_curriculum: the command curriculum is disabled in this config");
    if (rows != 2 * h->cfg.cur_lin_levels + 1 || cols != 2 * h->cfg.cur_ang_levels + 1)
        return set_error(B200_ERR_ARG, "b200_t1_bind_curriculum: grid shape must be [2 lin_vel_levels + 1, 2 ang_vel_levels + 1]");
    h->cur_prob = prob;
    return B200_OK;
}
int b200_t1_inject_rng(B200T1Handle* h, const uint32_t* table) {
    if (!h) return set_error(B200_ERR_ARG, "null handle");
    h->inject = table;
    return B200_OK;
}
int b200_t1_rng_slots(void) { return B200_RNG_SLOTS; }
int b200_t1_num_envs(const B200T1Handle* h) { return h ? h->num_envs : B200_ERR_ARG; }

#define NEED_STATE(h)                                                                  \
    if (!(h)) return set_error(B200_ERR_ARG, "null handle");                           \
    if (!(h)->fstate || !(h)->istate) return set_error(B200_ERR_STATE, "state not bound"); \
    if ((h)->cfg.curriculum && !(h)->cur_prob) return set_error(B200_ERR_STATE, "command curriculum enabled but no grid bound (b200_t1_bind_curriculum)")

static void launch_curriculum_apply(B200T1Handle* h, cudaStream_t st) {
    k_curriculum_apply<<<1, 512, 0, st>>>(h->cur_prob, h->cur_count, h->cur_cdf, h->cur_cells, h->cfg.cur_update_rate);
    g_launches += 1;
}

int b200_t1_init_params(B200T1Handle* h, int env_index_base, int total_envs, void* stream) {
    NEED_STATE(h);
    h->env_base = env_index_base;
    h->total_envs = total_envs;
    cudaStream_t st = (cudaStream_t)stream;
    k_init_params<<<(h->num_envs + 127) / 128, 128, 0, st>>>(make_view(h), h->model, h->cfg);
    g_launches += 1;
    return launch_status("k_init_params");
}

int b200_t1_reset(B200T1Handle* h, float* obs, float* priv, void* stream) {
    NEED_STATE(h);
    if (!obs || !priv) return set_error(B200_ERR_ARG, "b200_t1_reset: null output");
    cudaStream_t st = (cudaStream_t)stream;
    k_advance<<<1, 1, 0, st>>>(h->ctr_dev, -1, 0);
    const int grid = (h->num_envs + POST_BLOCK - 1) / POST_BLOCK;
    if (h->cfg.curriculum) {
        k_reset_all<1><<<grid, POST_BLOCK, 0, st>>>(make_view(h), h->model, h->cfg, make_terrain(h), h->ctr_dev, obs, priv);
        launch_curriculum_apply(h, st);
        k_reset_all<2><<<grid, POST_BLOCK, 0, st>>>(make_view(h), h->model, h->cfg, make_terrain(h), h->ctr_dev, obs, priv);
        g_launches += 1;
    } else {
        k_reset_all<0><<<grid, POST_BLOCK, 0, st>>>(make_view(h), h->model, h->cfg, make_terrain(h), h->ctr_dev, obs, priv);
    }
    g_launches += 2;
    return launch_status("k_reset_all");
}

int b200_t1_physics(B200T1Handle* h, const float* actions, int n_substeps, int apply_pd, float* qacc_out, void* stream) {
    NEED_STATE(h);
    if (!actions || n_substeps < 0) return set_error(B200_ERR_ARG, "b200_t1_physics: bad argument");
    cudaStream_t st = (cudaStream_t)stream;
    return launch_physics(h, actions, n_substeps, apply_pd, qacc_out, -1, 0, st);
}

static bool g_post_pair = getenv("B200_POST_PAIR") ? atoi(getenv("B200_POST_PAIR")) != 0 : true;   // 0: k_post<0> (one warp per 32 envs)
static int launch_post(B200T1Handle* h, float* obs, float* priv, float* rew, uint8_t* done, uint8_t* time_out,
                       float* rew_terms, int noise_on, cudaStream_t st) {
    const int grid = (h->num_envs + POST_BLOCK - 1) / POST_BLOCK;
    if (h->cfg.curriculum) {
        k_post<1><<<grid, POST_BLOCK, 0, st>>>(make_view(h), h->model, h->cfg, make_terrain(h), h->ctr_dev, noise_on, obs, priv, rew, done,
                                               rew_terms, h->ctr_dev + 2, h->stats_dev);
        launch_curriculum_apply(h, st);
        k_post<2><<<grid, POST_BLOCK, 0, st>>>(make_view(h), h->model, h->cfg, make_terrain(h), h->ctr_dev, noise_on, obs, priv, rew, done,
                                               rew_terms, h->ctr_dev + 2, h->stats_dev);
        g_launches += 1;
    } else if (g_post_pair) {
        k_post_pair<<<(h->num_envs + PAIR_ENVS - 1) / PAIR_ENVS, 2 * PAIR_ENVS, 0, st>>>(make_view(h), h->model, h->cfg, make_terrain(h), h->ctr_dev, noise_on,
                                                                                         obs, priv, rew, done, rew_terms, h->ctr_dev + 2, h->stats_dev);
    } else {
        k_post<0><<<grid, POST_BLOCK, 0, st>>>(make_view(h), h->model, h->cfg, make_terrain(h), h->ctr_dev, noise_on, obs, priv, rew, done,
                                               rew_terms, h->ctr_dev + 2, h->stats_dev);
    }
    k_finalize_timeouts<<<(h->num_envs + 255) / 256, 256, 0, st>>>(h->istate, h->num_envs, h->ctr_dev, h->ctr_dev + 2, time_out);
    g_launches += 2;
    return launch_status("k_post");
}

int b200_t1_post_physics(B200T1Handle* h, float* obs, float* priv, float* rew, uint8_t* done, uint8_t* time_out,
                         float* rew_terms, int64_t common_step, int noise_on, void* stream) {
    NEED_STATE(h);
    if (!obs || !priv || !rew || !done || !time_out) return set_error(B200_ERR_ARG, "b200_t1_post_physics: null output");
    cudaStream_t st = (cudaStream_t)stream;
    k_advance<<<1, 1, 0, st>>>(h->ctr_dev, (long long)common_step, 1);
    g_launches += 1;
    return launch_post(h, obs, priv, rew, done, time_out, rew_terms, noise_on, st);
}

int b200_t1_step(B200T1Handle* h, const float* actions, float* obs, float* priv, float* rew, uint8_t* done,
                 uint8_t* time_out, float* rew_terms, int64_t common_step, void* stream) {
    NEED_STATE(h);
    if (!actions || !obs || !priv || !rew || !done || !time_out) return set_error(B200_ERR_ARG, "b200_t1_step: null pointer");
    cudaStream_t st = (cudaStream_t)stream;
    const int rc = launch_physics(h, actions, h->cfg.decimation, 1, nullptr, (long long)common_step, 1, st);
    if (rc != B200_OK) return rc;
    return launch_post(h, obs, priv, rew, done, time_out, rew_terms, 1, st);
}

int b200_t1_episode_stats(B200T1Handle* h, double* sums_host, int64_t* count_host, void* stream) {
    if (!h || !sums_host || !count_host) return set_error(B200_ERR_ARG, "b200_t1_episode_stats: null pointer");
    cudaStream_t st = (cudaStream_t)stream;
    double tmp[B200_MAX_REW + 4];
    const int k = h->cfg.n_rew + 3;
    CUDA_TRY(cudaMemcpyAsync(tmp, h->stats_dev, k * sizeof(double), cudaMemcpyDeviceToHost, st));
    CUDA_TRY(cudaMemsetAsync(h->stats_dev, 0, k * sizeof(double), st));
    CUDA_TRY(cudaStreamSynchronize(st));
    for (int i = 0; i < k - 1; ++i) sums_host[i] = tmp[i];
    *count_host = (int64_t)(tmp[k - 1] + 0.5);
    return B200_OK;
}

int b200_t1_episode_stats_async(B200T1Handle* h, double* out_pinned, void* stream) {
    if (!h || !out_pinned) return set_error(B200_ERR_ARG, "b200_t1_episode_stats_async: null pointer");
    cudaStream_t st = (cudaStream_t)stream;
    const int k = h->cfg.n_rew + 3;
    CUDA_TRY(cudaMemcpyAsync(out_pinned, h->stats_dev, k * sizeof(double), cudaMemcpyDeviceToHost, st));
    CUDA_TRY(cudaMemsetAsync(h->stats_dev, 0, k * sizeof(double), st));
    return B200_OK;
}

int b200_terrain_heights(const B200T1Handle* h, const float* xy, int stride, int count, float* out, void* stream) {
    if (!h || !xy || !out || stride < 2 || count < 0) return set_error(B200_ERR_ARG, "b200_terrain_heights: bad argument");
    if (count == 0) return B200_OK;
    k_terrain_heights<<<(count + 255) / 256, 256, 0, (cudaStream_t)stream>>>(make_terrain(h), xy, stride, count, out);
    g_launches += 1;
    return launch_status("k_terrain_heights");
}

int b200_rng_fill(const B200T1Handle* h, uint64_t step, int purpose, int sub, int kind, float* out, void* stream) {
    if (!h || !out) return set_error(B200_ERR_ARG, "b200_rng_fill: null pointer");
    k_rng_fill<<<(h->num_envs + 255) / 256, 256, 0, (cudaStream_t)stream>>>(h->seed, h->env_base, h->num_envs, step, purpose, sub, kind, out);
    return launch_status("k_rng_fill");
}

/* current (rng step, common_step) as seen by the NEXT kernel in the stream; synchronises. Test helper. */
int b200_t1_counters(B200T1Handle* h, int64_t* rng_step, int64_t* common_step, void* stream) {
    if (!h) return set_error(B200_ERR_ARG, "null handle");
    long long tmp[2];
    CUDA_TRY(cudaMemcpyAsync(tmp, h->ctr_dev, sizeof tmp, cudaMemcpyDeviceToHost, (cudaStream_t)stream));
    CUDA_TRY(cudaStreamSynchronize((cudaStream_t)stream));
    if (rng_step) *rng_step = tmp[0];
    if (common_step) *common_step = tmp[1];
    return B200_OK;
}

}  // extern "C"
