// count_flops.cpp - TOOL: exact floating-point operation count of one physics tick (the numerator of k_physics' FP32 roofline).
// The tick is a template over the scalar type (booster_gym_b200/csrc/t1_dynamics.cuh); here it is instantiated with a scalar that
// counts every +, -, *, /, sqrt, rsqrt, sin/cos, abs/max/compare it performs while computing the same numbers in double.
//   python tools/count_flops.py     (builds this file with g++ and drives it through ctypes with the T1 model and two states)
// FMA convention of SURVEY 8(d): a multiply feeding an add counts as 2 FLOP either way, so  flop = adds + muls + divs + special.
#include <math.h>
#include <stdio.h>
#include <stddef.h>
#include <string.h>

struct Cnt;
static long long g_add = 0, g_mul = 0, g_div = 0, g_special = 0, g_cmp = 0;
struct Cnt {
    double v;
    Cnt() : v(0) {}
    Cnt(double x) : v(x) {}
    Cnt(float x) : v(x) {}
    Cnt(int x) : v(x) {}
    explicit operator double() const { return v; }
    explicit operator float() const { return (float)v; }
    Cnt operator-() const { return Cnt(-v); }
    Cnt& operator+=(const Cnt& o) { ++g_add; v += o.v; return *this; }
    Cnt& operator-=(const Cnt& o) { ++g_add; v -= o.v; return *this; }
    Cnt& operator*=(const Cnt& o) { ++g_mul; v *= o.v; return *this; }
};
static inline Cnt operator+(const Cnt& a, const Cnt& b) { ++g_add; return Cnt(a.v + b.v); }
static inline Cnt operator-(const Cnt& a, const Cnt& b) { ++g_add; return Cnt(a.v - b.v); }
static inline Cnt operator*(const Cnt& a, const Cnt& b) { ++g_mul; return Cnt(a.v * b.v); }
static inline Cnt operator/(const Cnt& a, const Cnt& b) { ++g_div; return Cnt(a.v / b.v); }
static inline bool operator<(const Cnt& a, const Cnt& b) { ++g_cmp; return a.v < b.v; }
static inline bool operator>(const Cnt& a, const Cnt& b) { ++g_cmp; return a.v > b.v; }
static inline bool operator<=(const Cnt& a, const Cnt& b) { ++g_cmp; return a.v <= b.v; }
static inline bool operator>=(const Cnt& a, const Cnt& b) { ++g_cmp; return a.v >= b.v; }
static inline bool operator==(const Cnt& a, const Cnt& b) { ++g_cmp; return a.v == b.v; }
static inline bool operator!=(const Cnt& a, const Cnt& b) { ++g_cmp; return a.v != b.v; }

// found by argument-dependent lookup from inside the templates (Cnt lives in the global namespace)
static inline void b_sincos(Cnt x, Cnt& s, Cnt& c) { g_special += 2; s = Cnt(sin(x.v)); c = Cnt(cos(x.v)); }
static inline Cnt b_sqrt(Cnt x) { ++g_special; return Cnt(sqrt(x.v)); }
static inline Cnt b_div(Cnt a, Cnt b) { ++g_div; return Cnt(a.v / b.v); }
static inline Cnt b_rsqrt(Cnt x) { ++g_special; return Cnt(1.0 / sqrt(x.v)); }
static inline Cnt b_abs(Cnt x) { ++g_cmp; return Cnt(fabs(x.v)); }
static inline Cnt b_max(Cnt a, Cnt b) { ++g_cmp; return Cnt(a.v > b.v ? a.v : b.v); }
static inline Cnt clamp01(Cnt x) { g_cmp += 2; return Cnt(x.v < 0 ? 0.0 : (x.v > 1 ? 1.0 : x.v)); }

#include "../booster_gym_b200/csrc/t1_env.cuh"

using namespace b200;

struct ModelC { B200_MODEL_FIELDS(Cnt) };   // the model in the counting scalar (same field list as B200T1ModelD)
static void model_to_cnt(const B200T1ModelD& d, ModelC& c) {
    // both structs are arrays of scalars followed by int32 fields in the same order: copy scalar by scalar
    const size_t n_real = offsetof(B200T1ModelD, axis) / sizeof(double);
    const double* src = reinterpret_cast<const double*>(&d);
    Cnt* dst = reinterpret_cast<Cnt*>(&c);
    for (size_t i = 0; i < n_real; ++i) dst[i] = Cnt(src[i]);
    memcpy(c.axis, d.axis, sizeof(d.axis));
    c.enable_contact = d.enable_contact; c.enable_limits = d.enable_limits; c.enable_body_contact = d.enable_body_contact;
    c.enable_self_contact = d.enable_self_contact; c.pad0 = 0;
}

struct FlatEnvD {  // oracle T1OEnv field order
    double pos[3], quat[4], vlin[3], wb[3], q[12], qd[12];
    double mass[B200_NB], com[B200_NB][3];
    double mu[2], kscale[2], cscale[2];
};

// one tick on the plane; counts[0..4] = adds, muls, divs, special (sqrt / rsqrt / sin / cos), compares (abs / max / min / <)
extern "C" int cf_tick(const B200T1ModelD* m, const FlatEnvD* e, const double* tau, long long* counts, double* foot_fn) {
    DynState<Cnt> s;
    DynParams<Cnt> p;
    for (int i = 0; i < 3; ++i) { s.pos[i] = e->pos[i]; s.vlin[i] = e->vlin[i]; s.wb[i] = e->wb[i]; }
    for (int i = 0; i < 4; ++i) s.quat[i] = e->quat[i];
    for (int i = 0; i < 12; ++i) { s.q[i] = e->q[i]; s.qd[i] = e->qd[i]; }
    for (int b = 0; b < B200_NB; ++b) { p.mass[b] = e->mass[b]; for (int k = 0; k < 3; ++k) p.com[b][k] = e->com[b][k]; }
    for (int k = 0; k < 2; ++k) { p.mu[k] = e->mu[k]; p.kscale[k] = e->kscale[k]; p.cscale[k] = e->cscale[k]; }
    Cnt t[12], pf[3], pt[3];
    for (int i = 0; i < 12; ++i) t[i] = tau[i];
    TerrainView tv{nullptr, 0, 0, 0, 0.1f, 0.005};
    tv.max_height = 0.0f;
    MLocal<Cnt> M;
    DynAux<Cnt> aux;
    static_assert(sizeof(Cnt) == sizeof(double), "Cnt is a plain double");
    ModelC mc;
    model_to_cnt(*m, mc);
    g_add = g_mul = g_div = g_special = g_cmp = 0;
    t1_tick<Cnt>(mc, p, s, t, pf, pt, tv, M, aux, true);
    counts[0] = g_add; counts[1] = g_mul; counts[2] = g_div; counts[3] = g_special; counts[4] = g_cmp;
    foot_fn[0] = aux.foot_fn[0].v; foot_fn[1] = aux.foot_fn[1].v;
    return 0;
}
