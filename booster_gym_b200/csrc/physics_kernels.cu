// physics_kernels.cu - K1: the decimated PD-torque physics loop of T1.step() (envs/t1.py:439-456 + gym.simulate).
//
// Mapping: TWO LANES PER ENVIRONMENT (one per leg; the base is carried redundantly) over structure-of-arrays state.
// The kernel runs the whole decimated loop (10 ticks of FK / CRBA / RNE / sparse LTDL / contact, t1_dynamics.cuh) in
// registers.  Model and config travel as __grid_constant__ kernel parameters (constant bank: the FFMA pipe reads them as
// operands instead of issuing loads).  Compiled WITH fma contraction (FP32-pipe / latency bound).
#include "env_handle.cuh"

using namespace b200;
namespace b200 { extern long long g_launches; }

// One environment = a pair of adjacent lanes (left / right leg); 16 environments per 32-thread CTA, so N = 4096 envs are
// 256 CTAs over the 148 SMs.  Everything lives in registers: the leg-parallel factorisation needs 21 + 36 + 21 matrix
// entries per lane (t1_dynamics.cuh), the pair exchanges 27 floats per tick by __shfl_xor_sync.
__global__ void __launch_bounds__(PHYS_BLOCK)
k_physics(EnvView v, const __grid_constant__ B200T1ModelF m, const __grid_constant__ B200T1Config c, TerrainView terr,
          const float* __restrict__ actions, int n_substeps, int apply_pd, float* __restrict__ qacc_out,
          long long* ctr, long long common_step, int advance) {
    if (advance && blockIdx.x == 0 && threadIdx.x == 0) {
        // start of a T1.step(): common_step_counter += 1 (envs/t1.py:477), new RNG epoch, clear the stale any-reset flag
        ctr[1] = (common_step >= 0) ? common_step : ctr[1] + 1;
        const long long s = ctr[0] + 1;
        ctr[0] = s;
        ctr[2 + ((s + 1) & 1)] = 0;
    }
    const int t = blockIdx.x * PHYS_BLOCK + threadIdx.x;
    const int side = t & 1;
    int e = t >> 1;
    const bool valid = e < v.n;
    if (!valid) e = v.n - 1;
    float act[6];
    const float2* a2 = reinterpret_cast<const float2*>(actions + (size_t)e * 12 + 6 * side);
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        const float2 q = a2[k];
        act[2 * k] = q.x; act[2 * k + 1] = q.y;
    }
    env_physics_pair(v, e, valid, side, m, c, terr, act, n_substeps, apply_pd, qacc_out);
}

namespace b200 {
int launch_physics(B200T1Handle* h, const float* actions, int n_substeps, int apply_pd, float* qacc_out,
                   long long common_step, int advance, cudaStream_t st) {
    k_physics<<<(2 * h->num_envs + PHYS_BLOCK - 1) / PHYS_BLOCK, PHYS_BLOCK, 0, st>>>(
        make_view(h), h->model, h->cfg, make_terrain(h), actions, n_substeps, apply_pd, qacc_out, h->ctr_dev,
        common_step, advance);
    g_launches += 1;
    return launch_status("k_physics");
}
}  // namespace b200
