// physics_kernels.cu - K1: the decimated PD-torque physics loop of T1.step() (envs/t1.py:439-456 + gym.simulate).
//
// Mapping: ONE THREAD PER ENVIRONMENT over structure-of-arrays state (coalesced rows).  The kernel runs the whole
// decimated loop (10 ticks of FK / CRBA / RNE / sparse LTDL / contact, t1_dynamics.cuh) with the 18x18 mass matrix of
// each thread in shared memory (interleaved by lane: bank-conflict free), one warp per CTA so that N = 4096 envs
// spread over 128 of the 148 SMs.  Model and config travel as __grid_constant__ kernel parameters (constant bank:
// the FFMA pipe reads them as operands instead of issuing loads).  Compiled WITH fma contraction (FP32-pipe bound).
#include "env_handle.cuh"

using namespace b200;
namespace b200 { extern long long g_launches; }

// ---- mass-matrix storage in shared memory: element idx of lane l at sm[idx * 32 + l] -------------------------------
struct MShared {
    float* base;
    __device__ __forceinline__ float& operator()(int i, int j) { return base[(i * (i + 1) / 2 + j) * PHYS_BLOCK]; }
};

__global__ void __launch_bounds__(PHYS_BLOCK)
k_physics(EnvView v, const __grid_constant__ B200T1ModelF m, const __grid_constant__ B200T1Config c, TerrainView terr,
          const float* __restrict__ actions, int n_substeps, int apply_pd, float* __restrict__ qacc_out,
          long long* ctr, long long common_step, int advance) {
    __shared__ float sM[PHYS_BLOCK * (B200_NV * (B200_NV + 1) / 2)];
    if (advance && blockIdx.x == 0 && threadIdx.x == 0) {
        // start of a T1.step(): common_step_counter += 1 (envs/t1.py:477), new RNG epoch, clear the stale any-reset flag
        ctr[1] = (common_step >= 0) ? common_step : ctr[1] + 1;
        const long long s = ctr[0] + 1;
        ctr[0] = s;
        ctr[2 + ((s + 1) & 1)] = 0;
    }
    const int e = blockIdx.x * PHYS_BLOCK + threadIdx.x;
    if (e >= v.n) return;
    MShared M;
    M.base = sM + threadIdx.x;
    float act[12];
    const float4* a4 = reinterpret_cast<const float4*>(actions + (size_t)e * 12);
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        const float4 t = a4[k];
        act[4 * k] = t.x; act[4 * k + 1] = t.y; act[4 * k + 2] = t.z; act[4 * k + 3] = t.w;
    }
    float qacc[B200_NV];
    env_physics(v, e, m, c, terr, act, n_substeps, apply_pd, M, qacc_out ? qacc : nullptr);
    if (qacc_out) {
#pragma unroll
        for (int i = 0; i < B200_NV; ++i) qacc_out[(size_t)i * v.n + e] = qacc[i];
    }
}


namespace b200 {
int launch_physics(B200T1Handle* h, const float* actions, int n_substeps, int apply_pd, float* qacc_out,
                   long long common_step, int advance, cudaStream_t st) {
    k_physics<<<(h->num_envs + PHYS_BLOCK - 1) / PHYS_BLOCK, PHYS_BLOCK, 0, st>>>(
        make_view(h), h->model, h->cfg, make_terrain(h), actions, n_substeps, apply_pd, qacc_out, h->ctr_dev,
        common_step, advance);
    g_launches += 1;
    return launch_status("k_physics");
}
}  // namespace b200
