// physics_port.cpp - CPU PORT of the decimated physics loop for bench.py's cpu_baseline / --impl reference legs.
// TEST / BENCH INFRASTRUCTURE ONLY: never loaded by the product package.
//
// The checker (physics_oracle.c) is deliberately a different, slow algorithm (dense Jacobians, O(nb nv^2)); timing it
// would flatter the GPU.  This file instead host-compiles the product's own O(nv) recursion (csrc/t1_dynamics.cuh:
// CRBA + RNE + sparse LTDL, the same operation count class as MuJoCo's mj_step for this tree) in FP64 at -O3, one env
// per OpenMP thread - the fairest CPU stand-in available while `mujoco` cannot be installed.
// Control law: envs/t1.py:444-456.
#include <math.h>
#include <string.h>

#include "../booster_gym_b200/csrc/t1_dynamics.cuh"
#include "../booster_gym_b200/csrc/terrain.cuh"

using namespace b200;

struct PortEnv {  // same field order as oracle T1OEnv
    double pos[3], quat[4], vlin[3], wb[3], q[12], qd[12];
    double mass[B200_NB], com[B200_NB][3];
    double mu[2], kscale[2], cscale[2];
};

extern "C" int t1p_env_physics(const B200T1ModelD* m, PortEnv* envs, int nenv, const double* actions, const double* default_q,
                               double action_scale, const double* kp, const double* kd, const double* fric,
                               const double* torque_limit, const int* delay, double* last_targets, int decimation,
                               double* torques_mean) {
    TerrainView terr{nullptr, 0, 0, 0, 0.1f, 0.005};
    terr.max_height = 0.0f;   // the plane: lets the shape culls of the tick work as they do on the device (a fair CPU baseline)
#pragma omp parallel for schedule(static)
    for (int n = 0; n < nenv; ++n) {
        DynState<double> s;
        DynParams<double> p;
        PortEnv* e = &envs[n];
        memcpy(s.pos, e->pos, 24); memcpy(s.quat, e->quat, 32); memcpy(s.vlin, e->vlin, 24); memcpy(s.wb, e->wb, 24);
        memcpy(s.q, e->q, 96); memcpy(s.qd, e->qd, 96);
        memcpy(p.mass, e->mass, sizeof p.mass); memcpy(p.com, e->com, sizeof p.com);
        memcpy(p.mu, e->mu, 16); memcpy(p.kscale, e->kscale, 16); memcpy(p.cscale, e->cscale, 16);
        MLocal<double> M;
        DynAux<double> aux;
        const double zero3[3] = {0, 0, 0};
        double tgt[12], tau[12], acc[12];
        for (int j = 0; j < 12; ++j) { tgt[j] = default_q[j] + action_scale * actions[12 * n + j]; acc[j] = 0; }
        for (int i = 0; i < decimation; ++i) {
            if (delay[n] == i)
                for (int j = 0; j < 12; ++j) last_targets[12 * n + j] = tgt[j];
            for (int j = 0; j < 12; ++j) {
                double t = kp[12 * n + j] * (last_targets[12 * n + j] - s.q[j]) - kd[12 * n + j] * s.qd[j];
                const double f = fmin(fric[12 * n + j], fabs(t));
                t -= (t > 0 ? f : (t < 0 ? -f : 0.0));
                t = fmax(-torque_limit[j], fmin(torque_limit[j], t));
                tau[j] = t;
                acc[j] += t;
            }
            t1_tick<double>(*m, p, s, tau, zero3, zero3, terr, M, aux, true);
        }
        memcpy(e->pos, s.pos, 24); memcpy(e->quat, s.quat, 32); memcpy(e->vlin, s.vlin, 24); memcpy(e->wb, s.wb, 24);
        memcpy(e->q, s.q, 96); memcpy(e->qd, s.qd, 96);
        for (int j = 0; j < 12; ++j) torques_mean[12 * n + j] = acc[j] / decimation;
    }
    return 0;
}
