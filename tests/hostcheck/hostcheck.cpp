// hostcheck.cpp - TEST-ONLY host build of the shared __host__ __device__ kernel bodies (csrc/*.cuh).
// Lets `pytest -m "not gpu"` exercise the exact arithmetic the CUDA kernels run, in float and double, against
// oracle/ without a GPU.  Never loaded by the product package (booster_gym_b200/_lib.py only loads libb200t1.so).
#include <string.h>

#include "../../booster_gym_b200/csrc/t1_env.cuh"

using namespace b200;

template <typename T> struct FlatEnv {  // mirrors oracle T1OEnv field order, in scalar type T
    T pos[3], quat[4], vlin[3], wb[3], q[12], qd[12];
    T mass[B200_NB], com[B200_NB][3];
    T mu[2], kscale[2], cscale[2];
};

template <typename T, typename Model>
static int tick_impl(const Model* m, FlatEnv<T>* e, const T* tau, const T* push_f, const T* push_t, const int16_t* hf,
                     int rows, int cols, int border_pixels, float hscale, double vscale, T* qacc, T* foot_fn,
                     int integrate, T* contact_out /*9, nullable: |F|^2 of hip-yaw, shank, foot per leg, then the trunk force*/) {
    DynState<T> s;
    DynParams<T> p;
    memcpy(s.pos, e->pos, sizeof(T) * 3); memcpy(s.quat, e->quat, sizeof(T) * 4);
    memcpy(s.vlin, e->vlin, sizeof(T) * 3); memcpy(s.wb, e->wb, sizeof(T) * 3);
    memcpy(s.q, e->q, sizeof(T) * 12); memcpy(s.qd, e->qd, sizeof(T) * 12);
    memcpy(p.mass, e->mass, sizeof(p.mass)); memcpy(p.com, e->com, sizeof(p.com));
    memcpy(p.mu, e->mu, sizeof(p.mu)); memcpy(p.kscale, e->kscale, sizeof(p.kscale)); memcpy(p.cscale, e->cscale, sizeof(p.cscale));
    TerrainView tv{hf, rows, cols, border_pixels, hscale, vscale};
    MLocal<T> M;
    DynAux<T> aux;
    t1_tick<T>(*m, p, s, tau, push_f, push_t, tv, M, aux, integrate != 0);
    memcpy(e->pos, s.pos, sizeof(T) * 3); memcpy(e->quat, s.quat, sizeof(T) * 4);
    memcpy(e->vlin, s.vlin, sizeof(T) * 3); memcpy(e->wb, s.wb, sizeof(T) * 3);
    memcpy(e->q, s.q, sizeof(T) * 12); memcpy(e->qd, s.qd, sizeof(T) * 12);
    if (qacc) memcpy(qacc, aux.qacc, sizeof(T) * B200_NV);
    if (foot_fn) { foot_fn[0] = aux.foot_fn[0]; foot_fn[1] = aux.foot_fn[1]; }
    if (contact_out) {
        for (int sd = 0; sd < 2; ++sd)
            for (int i = 0; i < 3; ++i) contact_out[3 * sd + i] = aux.body_f2[sd][i];
        for (int i = 0; i < 3; ++i) contact_out[6 + i] = aux.trunk_f[i];
    }
    return 0;
}

extern "C" {
int hc_tick_d(const B200T1ModelD* m, void* env, const double* tau, const double* push_f, const double* push_t,
              const int16_t* hf, int rows, int cols, int border_pixels, float hscale, double vscale, double* qacc,
              double* foot_fn, int integrate) {
    return tick_impl<double>(m, (FlatEnv<double>*)env, tau, push_f, push_t, hf, rows, cols, border_pixels, hscale, vscale, qacc, foot_fn, integrate, nullptr);
}
int hc_tick_d_contacts(const B200T1ModelD* m, void* env, const double* tau, const double* push_f, const double* push_t,
                       const int16_t* hf, int rows, int cols, int border_pixels, float hscale, double vscale, double* qacc,
                       double* foot_fn, int integrate, double* contact_out) {
    return tick_impl<double>(m, (FlatEnv<double>*)env, tau, push_f, push_t, hf, rows, cols, border_pixels, hscale, vscale, qacc, foot_fn, integrate, contact_out);
}
int hc_tick_f(const B200T1ModelF* m, void* env, const float* tau, const float* push_f, const float* push_t,
              const int16_t* hf, int rows, int cols, int border_pixels, float hscale, double vscale, float* qacc,
              float* foot_fn, int integrate) {
    return tick_impl<float>(m, (FlatEnv<float>*)env, tau, push_f, push_t, hf, rows, cols, border_pixels, hscale, vscale, qacc, foot_fn, integrate, nullptr);
}
float hc_terrain_height(const int16_t* hf, int rows, int cols, int border_pixels, float hscale, double vscale, float x, float y) {
    TerrainView tv{hf, rows, cols, border_pixels, hscale, vscale};
    return tv(x, y);
}
void hc_feet_d(const B200T1ModelD* m, void* env, double* pos, double* quat) {
    FlatEnv<double>* e = (FlatEnv<double>*)env;
    DynState<double> s;
    memcpy(s.pos, e->pos, 24); memcpy(s.quat, e->quat, 32); memcpy(s.q, e->q, 96);
    double fp[2][3], fq[2][4];
    t1_feet_fk<double>(*m, s, fp, fq);
    memcpy(pos, fp, sizeof fp); memcpy(quat, fq, sizeof fq);
}

// ---- the env bodies (t1_env.cuh) on host arrays: what k_post + k_finalize_timeouts / k_reset_all do, serially ----------
static TerrainView make_tv(const B200T1Config* c, const int16_t* hf, int rows, int cols) {
    TerrainView tv{(c->terrain_type == 0) ? nullptr : hf, rows, cols, c->border_pixels, c->horizontal_scale, c->vertical_scale};
    return tv;
}
int hc_env_post(const B200T1ModelF* m, const B200T1Config* c, const int16_t* hf, int rows, int cols, float* fstate,
                int32_t* istate, int n, const uint32_t* inject, long long common_step, unsigned long long step, int noise_on,
                float* obs, float* priv, float* rew, uint8_t* done, uint8_t* time_out_extras, float* rew_terms) {
    EnvView v{fstate, istate, n, 0, 42ull, inject};
    const TerrainView tv = make_tv(c, hf, rows, cols);
    int any = 0;
    for (int e = 0; e < n; ++e) {
        const StepOut o = env_post_physics(v, e, *m, *c, tv, common_step, step, noise_on, obs + (size_t)e * B200_NOBS,
                                           priv + (size_t)e * B200_NPRIV, rew_terms);
        rew[e] = o.rew;
        done[e] = (uint8_t)o.done;
        any |= o.done;
    }
    if (any)
        for (int e = 0; e < n; ++e) time_out_extras[e] = (uint8_t)istate[(size_t)I_time_out_buf * n + e];
    return any;
}
// what k_post_pair (env_kernels.cu) does with one env on two warps, serialised in the interleaving that is LEAST like env_post_physics:
// warp 1's right-foot refresh before warp 0's part A, then the snapshot, then warp 1's reset + tail BEFORE warp 0's reward terms.  If the
// parts were not independent in the way the kernel assumes, this order would show it (the kernel's barriers allow exactly these orders).
int hc_env_post_pair(const B200T1ModelF* m, const B200T1Config* c, const int16_t* hf, int rows, int cols, float* fstate,
                     int32_t* istate, int n, const uint32_t* inject, long long common_step, unsigned long long step, int noise_on,
                     float* obs, float* priv, float* rew, uint8_t* done, uint8_t* time_out_extras, float* rew_terms) {
    EnvView v{fstate, istate, n, 0, 42ull, inject};
    const TerrainView tv = make_tv(c, hf, rows, cols);
    int any = 0;
    for (int e = 0; e < n; ++e) {
        env_refresh_feet<B200T1ModelF, 2>(v, e, *m, tv);                                              // warp 1, before the first barrier
        const PostA a = env_post_a<B200T1ModelF, 1>(v, e, *m, *c, tv, common_step, step);             // warp 0
        RewardSnap snap;
        reward_snapshot(v, e, snap);                                                                  // warp 0, between the barriers
        if (a.reset) env_reset_one(v, e, *c, tv, step);                                               // warp 1 ...
        env_post_tail(v, e, *m, *c, tv, step, noise_on, obs + (size_t)e * B200_NOBS, priv + (size_t)e * B200_NPRIV);
        rew[e] = env_post_rewards(v, e, *c, snap, a.h_base, a.finite, rew_terms);                     // ... and only now warp 0's rewards
        done[e] = (uint8_t)(a.reset ? 1 : 0);
        any |= a.reset ? 1 : 0;
    }
    if (any)
        for (int e = 0; e < n; ++e) time_out_extras[e] = (uint8_t)istate[(size_t)I_time_out_buf * n + e];
    return any;
}
// the command-curriculum step (cfg.curriculum): what k_post<1>, k_curriculum_apply, k_post<2> do, serially
int hc_env_post_curriculum(const B200T1ModelF* m, const B200T1Config* c, float* fstate, int32_t* istate, int n, const uint32_t* inject,
                           long long common_step, unsigned long long step, int noise_on, float* prob, float* obs, float* priv,
                           float* rew, uint8_t* done, uint8_t* time_out_extras, float* rew_terms) {
    const int cells = (2 * c->cur_lin_levels + 1) * (2 * c->cur_ang_levels + 1);
    int32_t* count = new int32_t[cells]();
    float* cdf = new float[cells]();
    EnvView v{fstate, istate, n, 0, 42ull, inject, count, cdf};
    const TerrainView tv = make_tv(c, nullptr, 0, 0);
    int any = 0;
    for (int e = 0; e < n; ++e) {
        const StepOut o = env_post_physics<B200T1ModelF, 1>(v, e, *m, *c, tv, common_step, step, noise_on, obs + (size_t)e * B200_NOBS,
                                                            priv + (size_t)e * B200_NPRIV, rew_terms);
        rew[e] = o.rew;
        done[e] = (uint8_t)o.done;
        any |= o.done;
    }
    for (int i = 0; i < cells; ++i) curriculum_apply_cell(prob, count, i, c->cur_update_rate);
    curriculum_scan(prob, cdf, cells);
    for (int e = 0; e < n; ++e)
        env_post_physics<B200T1ModelF, 2>(v, e, *m, *c, tv, common_step, step, noise_on, obs + (size_t)e * B200_NOBS,
                                          priv + (size_t)e * B200_NPRIV, rew_terms);
    if (any)
        for (int e = 0; e < n; ++e) time_out_extras[e] = (uint8_t)istate[(size_t)I_time_out_buf * n + e];
    delete[] count;
    delete[] cdf;
    return any;
}
int hc_env_reset_all(const B200T1ModelF* m, const B200T1Config* c, const int16_t* hf, int rows, int cols, float* fstate,
                     int32_t* istate, int n, const uint32_t* inject, unsigned long long step, float* obs, float* priv) {
    EnvView v{fstate, istate, n, 0, 42ull, inject};
    const TerrainView tv = make_tv(c, hf, rows, cols);
    (void)m;
    for (int e = 0; e < n; ++e) {
        env_reset_one(v, e, *c, tv, step);
        if (istate[(size_t)I_episode_length_buf * n + e] == istate[(size_t)I_cmd_resample_time * n + e]) env_resample_command(v, e, *c, step);
        env_observations(v, e, *c, tv, step, 1, obs + (size_t)e * B200_NOBS, priv + (size_t)e * B200_NPRIV);
    }
    return 0;
}
void hc_sincos_bounded(const float* x, int n, float* s, float* c) {   // the device path of b_sincos(float), t1_dynamics.cuh
    for (int i = 0; i < n; ++i) sincos_bounded(x[i], s[i], c[i]);
}
void hc_philox(unsigned long long seed, unsigned int env, unsigned long long step, int purpose, int sub, unsigned int* words,
               float* uni, float* nrm) {
    const Philox4 p = rng_words(seed, env, step, purpose, sub);
    const Rand4 r = rand4(p);
    for (int i = 0; i < 4; ++i) { words[i] = p.w[i]; uni[i] = r.u[i]; nrm[i] = r.n[i]; }
}
}
