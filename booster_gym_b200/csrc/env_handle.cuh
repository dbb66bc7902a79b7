// env_handle.cuh - the opaque handle behind B200T1Handle and the views the kernels take (shared by the two env TUs:
// physics_kernels.cu is compiled with FMA contraction, env_kernels.cu with --fmad=false so that the bookkeeping /
// observation / reward arithmetic rounds exactly like the reference's separate torch ops).
#pragma once
#include <cuda_runtime.h>

#include "common.cuh"
#include "t1_env.cuh"

struct B200T1Handle {
    B200T1ModelF model;
    B200T1Config cfg;
    int num_envs, device;
    uint64_t seed;
    int env_base, total_envs;
    float* fstate;
    int32_t* istate;
    int16_t* hf_dev;
    int hf_rows, hf_cols;
    float hf_max;         // highest heightfield sample [m] (0 for the plane)
    long long* ctr_dev;   // [0] rng step, [1] common_step_counter, [2..3] any-reset flags (parity)
    double* stats_dev;    // [1 + n_rew + 1] episode sums, then count as double
    const uint32_t* inject;  // parity-test hook (b200_t1_inject_rng); null in production
    // command curriculum (cfg.curriculum): caller-owned grid, per-step success counts and running sums (library-owned)
    float* cur_prob;
    int32_t* cur_count;
    float* cur_cdf;
    int cur_cells;
};

namespace b200 {
inline TerrainView make_terrain(const B200T1Handle* h) {
    TerrainView t;
    t.hf = (h->cfg.terrain_type == 0) ? nullptr : h->hf_dev;
    t.rows = h->hf_rows;
    t.cols = h->hf_cols;
    t.border_pixels = h->cfg.border_pixels;
    t.horizontal_scale = h->cfg.horizontal_scale;
    t.vertical_scale = h->cfg.vertical_scale;
    t.max_height = (h->cfg.terrain_type == 0) ? 0.0f : h->hf_max;
    return t;
}
inline EnvView make_view(const B200T1Handle* h) {
    EnvView v;
    v.f = h->fstate;
    v.is = h->istate;
    v.n = h->num_envs;
    v.env_base = h->env_base;
    v.seed = h->seed;
    v.inject = h->inject;
    v.cur_count = h->cur_count;
    v.cur_cdf = h->cur_cdf;
    return v;
}
// defined in physics_kernels.cu
int launch_physics(B200T1Handle* h, const float* actions, int n_substeps, int apply_pd, float* qacc_out,
                   long long common_step, int advance, cudaStream_t st);
}  // namespace b200
