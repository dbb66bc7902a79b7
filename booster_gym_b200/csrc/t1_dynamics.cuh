// t1_dynamics.cuh - one physics tick (2 ms) of the Booster T1 tree for ONE environment.
//
// Replaces what `gym.simulate` does between envs/t1.py:450-455 (PhysX, closed source) and what `mujoco.mj_step`
// does at play_mujoco.py:756, on the model of resources/T1/T1_locomotion.xml:37-119.  Smooth dynamics follow the
// MuJoCo forward pipeline: generalised velocity [v_world, omega_body, qd], world-aligned
// spatial algebra about a common reference point (here the trunk origin instead of the subtree CoM, which keeps
// fp32 numbers small on an 80 m terrain), composite-rigid-body mass matrix, recursive Newton-Euler bias, sparse
// L^T D L factorisation in leaf-to-root order (left/right leg blocks never couple), semi-implicit Euler with an
// exact quaternion exponential.  Contact is NOT MuJoCo's convex solver: 8 sole corners against the heightfield with
// a linearly-implicit spring-damper normal force and a lagged, regularised Coulomb friction (DESIGN.md "contact"); the
// same law for the trunk box corners and the hip-yaw / shank cylinder rims, and a frictionless variant between the two
// legs' shank / foot capsules ("the rarely touching shapes" below).
//
// The code is a template over the scalar type and is __host__ __device__: the CUDA kernels instantiate it with
// float; tests/ build it for the host (float and double) to check the algebra against oracle/ without a GPU.
// The host build is test infrastructure only - the product library never runs it.
#pragma once
#include <math.h>
#include <stdint.h>

#include "../../include/b200_t1.h"

#if defined(__CUDACC__)
#define B200_HD __host__ __device__ __forceinline__
#define B200_COLD __host__ __device__ __noinline__
#else
#define B200_HD inline
#define B200_COLD inline __attribute__((noinline))
#endif

namespace b200 {

template <typename T> struct ModelOf;
template <> struct ModelOf<float> { typedef B200T1ModelF type; };
template <> struct ModelOf<double> { typedef B200T1ModelD type; };

// sin / cos for BOUNDED arguments (joint angles, half rotation angles: a few radians).  The tick is instruction-fetch bound and
// needs 7 of these: CUDA's sincosf inlines ~150 instructions each (Payne-Hanek slow path included); a 3-constant Cody-Waite
// reduction to [-pi/4, pi/4] (exact products for |k| < 2^13) + the cephes minimax polynomials take ~25, |error| < 1e-7 for
// |x| <= 8 (tests/test_kernel_bodies_host.py::test_bounded_sincos).
B200_HD void sincos_bounded(float x, float& s, float& c) {
    const float k = rintf(x * 0.636619772367581343f);
    float r = fmaf(-k, 1.5703125f, x);
    r = fmaf(-k, 4.837512969970703125e-4f, r);
    r = fmaf(-k, 7.54978995489188216e-8f, r);
    const float z = r * r;
    const float sp = fmaf(fmaf(fmaf(-1.9515295891e-4f, z, 8.3321608736e-3f), z, -1.6666654611e-1f) * z, r, r);
    const float cp = fmaf(fmaf(fmaf(2.443315711809948e-5f, z, -1.388731625493765e-3f), z, 4.166664568298827e-2f), z * z, fmaf(-0.5f, z, 1.0f));
    const int q = (int)k;
    const float ss = (q & 1) ? cp : sp, cc = (q & 1) ? sp : cp;
    s = (q & 2) ? -ss : ss;
    c = ((q + 1) & 2) ? -cc : cc;
}
B200_HD void b_sincos(float x, float& s, float& c) {
#if defined(__CUDA_ARCH__)
    sincos_bounded(x, s, c);
#else
    s = sinf(x); c = cosf(x);
#endif
}
B200_HD void b_sincos(double x, double& s, double& c) { s = sin(x); c = cos(x); }
B200_HD float b_sqrt(float x) { return sqrtf(x); }
B200_HD double b_sqrt(double x) { return sqrt(x); }
// division in the rarely executed shape code: approximate on the device (2 ulp), exact on the host
B200_HD float b_div(float a, float b) {
#if defined(__CUDA_ARCH__)
    // one MUFU.RCP + one FMUL (__fdividef adds a 4-instruction rescue for |b| > 2^126, which no divisor here reaches)
    float r;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(b));
    return a * r;
#else
    return a / b;
#endif
}
B200_HD double b_div(double a, double b) { return a / b; }
B200_HD float b_rsqrt(float x) {
#if defined(__CUDA_ARCH__)
    float r;   // one MUFU.RSQ (rsqrtf adds denormal scaling; every argument here is a sum of squares far above 1e-38)
    asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
#else
    return 1.0f / sqrtf(x);
#endif
}
B200_HD double b_rsqrt(double x) { return 1.0 / sqrt(x); }
B200_HD float b_abs(float x) { return fabsf(x); }
B200_HD double b_abs(double x) { return fabs(x); }
B200_HD float b_max(float a, float b) { return fmaxf(a, b); }
B200_HD double b_max(double a, double b) { return fmax(a, b); }

// ---- small vector helpers (all fully unrolled by construction) -------------------------------------------------
template <typename T> B200_HD void cross3(const T* a, const T* b, T* o) {
    T x = a[1] * b[2] - a[2] * b[1], y = a[2] * b[0] - a[0] * b[2], z = a[0] * b[1] - a[1] * b[0];
    o[0] = x; o[1] = y; o[2] = z;
}
template <typename T> B200_HD T dot3(const T* a, const T* b) { return a[0] * b[0] + a[1] * b[1] + a[2] * b[2]; }

// quaternion xyzw -> rotation matrix (row-major R[r][c]), body -> world
template <typename T> B200_HD void quat_to_mat(const T* q, T R[3][3]) {
    T x = q[0], y = q[1], z = q[2], w = q[3];
    R[0][0] = 1 - 2 * (y * y + z * z); R[0][1] = 2 * (x * y - w * z); R[0][2] = 2 * (x * z + w * y);
    R[1][0] = 2 * (x * y + w * z); R[1][1] = 1 - 2 * (x * x + z * z); R[1][2] = 2 * (y * z - w * x);
    R[2][0] = 2 * (x * z - w * y); R[2][1] = 2 * (y * z + w * x); R[2][2] = 1 - 2 * (x * x + y * y);
}

// rotation matrix -> quaternion xyzw (w >= 0 branch-stable Shepperd form)
template <typename T> B200_HD void mat_to_quat(const T R[3][3], T* q) {
    // the branch's largest component is sqrt(x) / 2, the others (sums of off-diagonal pairs) / (2 sqrt(x)): one rsqrt, no division
    T tr = R[0][0] + R[1][1] + R[2][2];
    if (tr > 0) {
        const T x = tr + 1, rs = b_rsqrt(x), h = T(0.5) * rs;
        q[3] = T(0.5) * x * rs; q[0] = (R[2][1] - R[1][2]) * h; q[1] = (R[0][2] - R[2][0]) * h; q[2] = (R[1][0] - R[0][1]) * h;
    } else if (R[0][0] > R[1][1] && R[0][0] > R[2][2]) {
        const T x = 1 + R[0][0] - R[1][1] - R[2][2], rs = b_rsqrt(x), h = T(0.5) * rs;
        q[3] = (R[2][1] - R[1][2]) * h; q[0] = T(0.5) * x * rs; q[1] = (R[0][1] + R[1][0]) * h; q[2] = (R[0][2] + R[2][0]) * h;
    } else if (R[1][1] > R[2][2]) {
        const T x = 1 + R[1][1] - R[0][0] - R[2][2], rs = b_rsqrt(x), h = T(0.5) * rs;
        q[3] = (R[0][2] - R[2][0]) * h; q[0] = (R[0][1] + R[1][0]) * h; q[1] = T(0.5) * x * rs; q[2] = (R[1][2] + R[2][1]) * h;
    } else {
        const T x = 1 + R[2][2] - R[0][0] - R[1][1], rs = b_rsqrt(x), h = T(0.5) * rs;
        q[3] = (R[1][0] - R[0][1]) * h; q[0] = (R[0][2] + R[2][0]) * h; q[1] = (R[1][2] + R[2][1]) * h; q[2] = T(0.5) * x * rs;
    }
}

// R <- R * Rot(axis, angle): only two columns change
template <typename T> B200_HD void rotate_about_axis(T R[3][3], int axis, T angle) {
    T s, c;
    b_sincos(angle, s, c);
    const int i = (axis + 1) % 3, j = (axis + 2) % 3;  // Rot(axis): col_i' = c col_i + s col_j ; col_j' = -s col_i + c col_j
#pragma unroll
    for (int r = 0; r < 3; ++r) {
        T a = R[r][i], b = R[r][j];
        R[r][i] = c * a + s * b;
        R[r][j] = c * b - s * a;
    }
}

// ---- spatial inertia about the reference point, world axes: (m, h = m c, I) with I = xx yy zz xy xz yz --------------
template <typename T> struct SpInertia {
    T m, h[3], I[6];
};
template <typename T> B200_HD void si_mul(const SpInertia<T>& s, const T* w, const T* v, T* n, T* f) {
    // [n; f] = I_sp [w; v]:  n = I w + h x v ,  f = m v - h x w
    T hv[3], hw[3];
    cross3(s.h, v, hv);
    cross3(s.h, w, hw);
    n[0] = s.I[0] * w[0] + s.I[3] * w[1] + s.I[4] * w[2] + hv[0];
    n[1] = s.I[3] * w[0] + s.I[1] * w[1] + s.I[5] * w[2] + hv[1];
    n[2] = s.I[4] * w[0] + s.I[5] * w[1] + s.I[2] * w[2] + hv[2];
    f[0] = s.m * v[0] - hw[0];
    f[1] = s.m * v[1] - hw[1];
    f[2] = s.m * v[2] - hw[2];
}
template <typename T> B200_HD void si_add(SpInertia<T>& a, const SpInertia<T>& b) {
    a.m += b.m;
#pragma unroll
    for (int i = 0; i < 3; ++i) a.h[i] += b.h[i];
#pragma unroll
    for (int i = 0; i < 6; ++i) a.I[i] += b.I[i];
}
// body inertia (body frame, about CoM, scaled) -> spatial inertia about the reference point in world axes
template <typename T>
B200_HD void si_from_body(const T R[3][3], const T* x /*body origin rel. ref*/, const T* ipos, const T* Ib, T mass,
                          T iscale, SpInertia<T>& o) {
    T c[3];
#pragma unroll
    for (int r = 0; r < 3; ++r) c[r] = x[r] + R[r][0] * ipos[0] + R[r][1] * ipos[1] + R[r][2] * ipos[2];
    // A = R * Ib (Ib symmetric: xx yy zz xy xz yz)
    T A[3][3];
#pragma unroll
    for (int r = 0; r < 3; ++r) {
        A[r][0] = R[r][0] * Ib[0] + R[r][1] * Ib[3] + R[r][2] * Ib[4];
        A[r][1] = R[r][0] * Ib[3] + R[r][1] * Ib[1] + R[r][2] * Ib[5];
        A[r][2] = R[r][0] * Ib[4] + R[r][1] * Ib[5] + R[r][2] * Ib[2];
    }
    T Ixx = A[0][0] * R[0][0] + A[0][1] * R[0][1] + A[0][2] * R[0][2];
    T Iyy = A[1][0] * R[1][0] + A[1][1] * R[1][1] + A[1][2] * R[1][2];
    T Izz = A[2][0] * R[2][0] + A[2][1] * R[2][1] + A[2][2] * R[2][2];
    T Ixy = A[0][0] * R[1][0] + A[0][1] * R[1][1] + A[0][2] * R[1][2];
    T Ixz = A[0][0] * R[2][0] + A[0][1] * R[2][1] + A[0][2] * R[2][2];
    T Iyz = A[1][0] * R[2][0] + A[1][1] * R[2][1] + A[1][2] * R[2][2];
    o.m = mass;
    o.h[0] = mass * c[0]; o.h[1] = mass * c[1]; o.h[2] = mass * c[2];
    o.I[0] = iscale * Ixx + mass * (c[1] * c[1] + c[2] * c[2]);
    o.I[1] = iscale * Iyy + mass * (c[0] * c[0] + c[2] * c[2]);
    o.I[2] = iscale * Izz + mass * (c[0] * c[0] + c[1] * c[1]);
    o.I[3] = iscale * Ixy - mass * c[0] * c[1];
    o.I[4] = iscale * Ixz - mass * c[0] * c[2];
    o.I[5] = iscale * Iyz - mass * c[1] * c[2];
}


// ---- penalty contact of ONE point of a body against the ground (normal = +z) ---------------------------------------------
// Linearly-implicit normal spring-damper + lagged, velocity-regularised Coulomb friction (DESIGN.md "contact").  p = the point
// relative to the reference point, (w, v) = spatial velocity of the body about the reference point.  Adds the explicit force
// to the wrench (Wn about the reference point, Wf) and dt * J^T D J to K (6x6 symmetric, lower-tri, order [ang; lin],
// entry i stored at K[i * KS]: KS = 1 for a register array, KS = block size for the per-thread shared-memory slices).
template <int KS, typename T>
B200_HD bool contact_ground_point(const T* p, T depth, const T* w, const T* v, T kn, T cn, T dn, T mu, T dt, T vstick, T* K, T* Wn,
                                  T* Wf, T& fn_sum) {
    T wxp[3];
    cross3(w, p, wxp);
    const T vc[3] = {v[0] + wxp[0], v[1] + wxp[1], v[2] + wxp[2]};
    const T fn0 = kn * depth - cn * vc[2];
    if (!(fn0 > 0)) return false;
    fn_sum += fn0;
    const T vt2 = vc[0] * vc[0] + vc[1] * vc[1];
    const T dtan = mu * fn0 * ((vt2 > vstick * vstick) ? b_rsqrt(vt2) : b_div(T(1), vstick));   // mu fn0 / max(|v_t|, vstick)
    const T Fe[3] = {-dtan * vc[0], -dtan * vc[1], kn * depth - dn * vc[2]};
    T pxF[3];
    cross3(p, Fe, pxF);
#pragma unroll
    for (int r = 0; r < 3; ++r) { Wn[r] += pxF[r]; Wf[r] += Fe[r]; }
    // K += dt * sum_d D_d r_d^T r_d with rows r_x = [0, pz, -py | 1 0 0], r_y = [-pz, 0, px | 0 1 0],
    // r_z = [py, -px, 0 | 0 0 1]   (point velocity = v + w x p)
    const T Dx = dt * dtan, Dy = dt * dtan, Dz = dt * dn;
    K[0 * KS] += Dy * p[2] * p[2] + Dz * p[1] * p[1];   // (0,0)
    K[1 * KS] += -Dz * p[0] * p[1];                     // (1,0)
    K[2 * KS] += Dx * p[2] * p[2] + Dz * p[0] * p[0];   // (1,1)
    K[3 * KS] += -Dy * p[0] * p[2];                     // (2,0)
    K[4 * KS] += -Dx * p[1] * p[2];                     // (2,1)
    K[5 * KS] += Dx * p[1] * p[1] + Dy * p[0] * p[0];   // (2,2)
    K[7 * KS] += Dx * p[2];                             // (3,1)
    K[8 * KS] += -Dx * p[1];                            // (3,2)
    K[9 * KS] += Dx;                                    // (3,3)
    K[10 * KS] += -Dy * p[2];                           // (4,0)
    K[12 * KS] += Dy * p[0];                            // (4,2)
    K[14 * KS] += Dy;                                   // (4,4)
    K[15 * KS] += Dz * p[1];                            // (5,0)
    K[16 * KS] += -Dz * p[0];                           // (5,1)
    K[20 * KS] += Dz;                                   // (5,5)
    return true;
}

// ---- the rarely touching shapes: trunk box, hip-yaw and shank cylinders, leg-leg capsules -----------------------
// k_physics is one warp per CTA at 255 registers and streams its code from L2: shape code placed inline in the tick cost
// 10 % even when it never ran.  So the tick only PUBLISHES the few frames the shapes need into a per-thread scratch (shared
// memory on the device, stride KS = block size; a plain array with KS = 1 on the host) and applies exact conservative culls
// (lowest point of a shape above the highest terrain sample; the two legs on their own sides of the trunk's sagittal
// plane).  When a ground cull fails it calls shapes_eval() OUT OF LINE, which reads the scratch and leaves one slot per body
// (slot j = 0 trunk share, 1 hip-yaw link, 2 shank); the leg-leg capsule pairs read the partner leg lane's scratch inline
// (t1_leg_phase1); the shank's contributions of both kinds meet in slot 2, the foot's stay in registers:
//   Kx[(27 j + i) KS], i < 21: implicit contact matrix (6x6 lower-tri, [ang; lin]);  i = 21..23: wrench moment about the
//   reference point;  i = 24..26: net contact force.
// Published geometry, at B200_KX_GEOM + :  0..5 shank capsule axis ends (+z, -z; relative to the trunk origin, world axes),
// 6..11 shank spatial velocity (w, v) about the trunk origin, 12..17 foot capsule axis ends, 18..23 foot spatial velocity,
// 24..26 hip-yaw cylinder centre, 27..29 its axis, 30..35 hip-yaw link spatial velocity, 36 `inward` (publish_shank);
// 37..45 trunk rotation, 46..48 trunk position, 49..54 trunk spatial velocity (written only ahead of a trunk evaluation).
#define B200_KX_SLOT 27
#define B200_KX_GEOM (3 * B200_KX_SLOT)
#define B200_KX_SIZE (B200_KX_GEOM + 55)
#if defined(__CUDA_ARCH__)
#define B200_SYNCWARP() __syncwarp()
#else
#define B200_SYNCWARP()
#endif

// accumulators of one scratch slot while shapes_eval runs (K itself accumulates in the scratch)
template <typename T> struct ShapeAcc {
    T Wn[3], Wf[3];
    bool init, active;
};
template <int KS, typename T> B200_HD void shape_slot_begin(ShapeAcc<T>& acc, T* K) {
    if (acc.init) return;
    acc.init = true;
#pragma unroll
    for (int i = 0; i < 21; ++i) K[i * KS] = 0;
}

// One collision cylinder (centre c relative to the reference point, unit axis a; resources/T1/T1_locomotion.xml:66,71,99,104)
// against the ground: the lowest rim point of each end cap is a contact point (the deepest points of a cylinder over a
// flat patch unless it stands on a cap).
template <int KS, typename T, typename Model, typename Terr>
B200_HD void cylinder_ground(const Model& m, int ci, const T* c, const T* a, const T* pos, const T* w, const T* v, const Terr& terr,
                             T* K, ShapeAcc<T>& acc) {
    const T rad = m.cyl_radius[ci], half = m.cyl_half[ci];
    const T n2 = b_max(T(1) - a[2] * a[2], T(0));
    shape_slot_begin<KS>(acc, K);
    // lowest rim point of a cap = cap centre - rad * d / |d| with d = z - (z.a) a  (|d|^2 = 1 - a_z^2)
    const T sc = (n2 > T(1e-12)) ? rad / b_sqrt(n2) : T(0);
    const T off[3] = {sc * a[2] * a[0], sc * a[2] * a[1], -sc * n2};
    const T kn = m.contact_k, cn = m.contact_c, dn = cn + m.dt * kn;
    T fsum = 0;
#pragma unroll
    for (int e = 0; e < 2; ++e) {
        const T hs = e ? -half : half;
        const T p[3] = {c[0] + hs * a[0] + off[0], c[1] + hs * a[1] + off[1], c[2] + hs * a[2] + off[2]};
        const T ground = (T)terr((float)(pos[0] + p[0]), (float)(pos[1] + p[1]));
        const T depth = ground - (pos[2] + p[2]);
        if (depth > 0 && contact_ground_point<KS>(p, depth, w, v, kn, cn, dn, m.body_mu, m.dt, m.stiction_vel, K, acc.Wn, acc.Wf, fsum))
            acc.active = true;
    }
}

// The trunk box (resources/T1/T1_locomotion.xml:42): the 4 corners on the side of this leg lane (side 0: +y, 1: -y).
template <int KS, typename T, typename Model, typename Terr>
B200_HD void trunk_ground(const Model& m, int side, const T R0[3][3], const T* pos, const T* w, const T* v, const Terr& terr, T* K,
                          ShapeAcc<T>& acc) {
    const T* bp = m.trunk_box_pos;
    const T* bh = m.trunk_box_half;
    const T ly = bp[1] + (side == 0 ? bh[1] : -bh[1]);
    shape_slot_begin<KS>(acc, K);
    const T kn = m.contact_k, cn = m.contact_c, dn = cn + m.dt * kn;
    T fsum = 0;
#pragma unroll
    for (int c = 0; c < 4; ++c) {
        const T lx = bp[0] + ((c & 1) ? -bh[0] : bh[0]), lz = bp[2] + ((c & 2) ? -bh[2] : bh[2]);
        T p[3];
#pragma unroll
        for (int r = 0; r < 3; ++r) p[r] = R0[r][0] * lx + R0[r][1] * ly + R0[r][2] * lz;
        const T ground = (T)terr((float)(pos[0] + p[0]), (float)(pos[1] + p[1]));
        const T depth = ground - (pos[2] + p[2]);
        if (depth > 0 && contact_ground_point<KS>(p, depth, w, v, kn, cn, dn, m.body_mu, m.dt, m.stiction_vel, K, acc.Wn, acc.Wf, fsum))
            acc.active = true;
    }
}

// ---- leg-leg contacts (asset.self_collisions: 0 = enabled, envs/T1.yaml:69) -------------------------------------------------
// The shank cylinder and the foot box of each leg are treated as capsules (segment + radius).  Every (left, right) pair of
// capsules closer than the sum of the radii pushes the two bodies apart along the line between the closest axis points with
// a linearly-implicit spring-damper on the RELATIVE normal velocity; each leg lane is implicit in its own velocity only (the
// partner's velocity enters explicitly), so the legs stay uncoupled in the mass matrix and the forces are equal and opposite.
// closest points of two segments P1 + s d1, P2 + t d2 (s, t in [0, 1]); both segments have positive length
B200_HD float clamp01(float x) {
#if defined(__CUDA_ARCH__)
    return __saturatef(x);
#else
    return x < 0 ? 0.0f : (x > 1 ? 1.0f : x);
#endif
}
B200_HD double clamp01(double x) { return x < 0 ? 0.0 : (x > 1 ? 1.0 : x); }
template <typename T> B200_HD void segment_closest(const T* P1, const T* d1, const T* P2, const T* d2, T& s, T& t) {
    // branch-free (selects only): the tick evaluates two of these on every step and lets the scheduler interleave them
    const T r[3] = {P1[0] - P2[0], P1[1] - P2[1], P1[2] - P2[2]};
    const T a = dot3(d1, d1), e = dot3(d2, d2), f = dot3(d2, r), c = dot3(d1, r), b = dot3(d1, d2);
    const T denom = a * e - b * b;
    const T ia = b_div(T(1), a), ie = b_div(T(1), e);
    const T s0 = (denom > T(1e-12) * a * e) ? clamp01(b_div(b * f - c * e, denom)) : T(0);
    const T t0 = (b * s0 + f) * ie;
    const T s_lo = clamp01(-c * ia), s_hi = clamp01((b - c) * ia);
    s = (t0 < 0) ? s_lo : ((t0 > 1) ? s_hi : s0);
    t = clamp01(t0);
}

// Bounding spheres of MY capsule i and the partner's capsule j (0 shank, 1 foot) from the published geometry (G mine, Gp the
// partner's: axis ends at 12 i + 0..5): false = too far apart to touch.
template <int KS, typename T, typename Model>
B200_HD bool capsule_spheres_touch(const Model& m, int i, int j, const T* G, const T* Gp) {
    T cc[3], la = 0, lb = 0;
#pragma unroll
    for (int r = 0; r < 3; ++r) {
        const T a0 = G[(12 * i + r) * KS], a1 = G[(12 * i + 3 + r) * KS], b0 = Gp[(12 * j + r) * KS], b1 = Gp[(12 * j + 3 + r) * KS];
        cc[r] = T(0.5) * ((a0 + a1) - (b0 + b1));
        la += (a1 - a0) * (a1 - a0);
        lb += (b1 - b0) * (b1 - b0);
    }
    const T reach = T(0.5) * (b_sqrt(la) + b_sqrt(lb)) + ((i == 0) ? m.cyl_radius[1] : m.foot_cap_radius) + ((j == 0) ? m.cyl_radius[1] : m.foot_cap_radius);
    return !(dot3(cc, cc) > reach * reach);
}

// One (mine i, partner j) capsule pair: true if they press on each other; then pm = my axis point relative to the reference
// point, n = unit normal from the partner's axis point to mine, fmag = the explicit force along n on my body [N].  The link
// spatial velocities are at 12 i + 6..11 of the published geometry.
template <int KS, typename T, typename Model>
B200_HD bool capsule_pair(const Model& m, int side, int i, int j, const T* G, const T* Gp, T* pm, T* n, T& fmag) {
    const T rsum = ((i == 0) ? m.cyl_radius[1] : m.foot_cap_radius) + ((j == 0) ? m.cyl_radius[1] : m.foot_cap_radius);
    const T ks = m.self_k, cs = m.self_c, dn = cs + m.dt * ks;
    T pa[3], dm[3], pb[3], dother[3];
#pragma unroll
    for (int r = 0; r < 3; ++r) {
        pa[r] = G[(12 * i + r) * KS];
        dm[r] = G[(12 * i + 3 + r) * KS] - pa[r];
        pb[r] = Gp[(12 * j + r) * KS];
        dother[r] = Gp[(12 * j + 3 + r) * KS] - pb[r];
    }
    // always evaluate (left segment, right segment) so that both lanes of an env pick the same closest points
    T sl, sr;
    if (side == 0) segment_closest(pa, dm, pb, dother, sl, sr);
    else segment_closest(pb, dother, pa, dm, sl, sr);
    const T um = (side == 0) ? sl : sr, uo = (side == 0) ? sr : sl;
    T po[3], d[3];
#pragma unroll
    for (int r = 0; r < 3; ++r) {
        pm[r] = pa[r] + um * dm[r];
        po[r] = pb[r] + uo * dother[r];
        d[r] = pm[r] - po[r];
    }
    const T d2 = dot3(d, d);
    const bool touching = (d2 < rsum * rsum) && (d2 > T(1e-12));
    const T inv = b_rsqrt(b_max(d2, T(1e-12)));
    const T dist = d2 * inv;
    const T depth = rsum - dist;
#pragma unroll
    for (int r = 0; r < 3; ++r) n[r] = d[r] * inv;
    // relative velocity of the two axis points (rigid-body velocity v + w x p of each link)
    T wm[3], wo[3], t1[3], t2[3], vr[3];
#pragma unroll
    for (int r = 0; r < 3; ++r) { wm[r] = G[(12 * i + 6 + r) * KS]; wo[r] = Gp[(12 * j + 6 + r) * KS]; }
    cross3(wm, pm, t1);
    cross3(wo, po, t2);
#pragma unroll
    for (int r = 0; r < 3; ++r) vr[r] = (G[(12 * i + 9 + r) * KS] + t1[r]) - (Gp[(12 * j + 9 + r) * KS] + t2[r]);
    const T vn = dot3(vr, n);
    fmag = ks * depth - dn * vn;
    return touching && (ks * depth - cs * vn > 0);
}
// fold an active pair into (K, Wn, Wf):  K += dt dn r r^T with r = [pm x n ; n]  (normal velocity of my axis point = r . [w; v])
template <int KS, typename T, typename Model>
B200_HD void capsule_pair_apply(const Model& m, const T* pm, const T* n, T fmag, T* K, T* Wn, T* Wf) {
    const T Fe[3] = {fmag * n[0], fmag * n[1], fmag * n[2]};
    T pxF[3], pxn[3];
    cross3(pm, Fe, pxF);
    cross3(pm, n, pxn);
#pragma unroll
    for (int r = 0; r < 3; ++r) { Wn[r] += pxF[r]; Wf[r] += Fe[r]; }
    const T rr[6] = {pxn[0], pxn[1], pxn[2], n[0], n[1], n[2]};
    const T D = m.dt * (m.self_c + m.dt * m.self_k);
#pragma unroll
    for (int a = 0; a < 6; ++a)
#pragma unroll
        for (int b = 0; b <= a; ++b) K[(a * (a + 1) / 2 + b) * KS] += D * rr[a] * rr[b];
}

// What the tick calls when a cull fails.  `what`: bit 0 trunk box, 1 hip-yaw cylinder, 2 shank cylinder (all against the
// ground).  Kx = this lane's scratch, Kp = the partner leg lane's (same env).  Returns a mask: bit j =
// slot j holds an active contact.
template <int KS, typename T, typename Model, typename Terr>
B200_COLD int shapes_eval(const Model& m, int side, int what, const Terr terr, T* Kx, const T* Kp) {
    ShapeAcc<T> acc[3];
#pragma unroll
    for (int j = 0; j < 3; ++j) {
        acc[j].init = false; acc[j].active = false;
#pragma unroll
        for (int r = 0; r < 3; ++r) { acc[j].Wn[r] = 0; acc[j].Wf[r] = 0; }
    }
    const T* G = Kx + B200_KX_GEOM * KS;
    T pos[3];
#pragma unroll
    for (int r = 0; r < 3; ++r) pos[r] = G[(46 + r) * KS];
    if (what & 1) {
        T R0[3][3], w[3], v[3];
#pragma unroll
        for (int r = 0; r < 3; ++r) {
#pragma unroll
            for (int c = 0; c < 3; ++c) R0[r][c] = G[(37 + 3 * r + c) * KS];
            w[r] = G[(49 + r) * KS]; v[r] = G[(52 + r) * KS];
        }
        trunk_ground<KS>(m, side, R0, pos, w, v, terr, Kx, acc[0]);
    }
    if (what & 2) {
        T c[3], a[3], w[3], v[3];
#pragma unroll
        for (int r = 0; r < 3; ++r) { c[r] = G[(24 + r) * KS]; a[r] = G[(27 + r) * KS]; w[r] = G[(30 + r) * KS]; v[r] = G[(33 + r) * KS]; }
        cylinder_ground<KS>(m, 0, c, a, pos, w, v, terr, Kx + B200_KX_SLOT * KS, acc[1]);
    }
    if (what & 4) {
        T c[3], a[3], w[3], v[3];
        const T ih = T(0.5) / m.cyl_half[1];
#pragma unroll
        for (int r = 0; r < 3; ++r) {
            const T p0 = G[r * KS], p1 = G[(3 + r) * KS];
            c[r] = T(0.5) * (p0 + p1); a[r] = (p0 - p1) * ih;
            w[r] = G[(6 + r) * KS]; v[r] = G[(9 + r) * KS];
        }
        cylinder_ground<KS>(m, 1, c, a, pos, w, v, terr, Kx + B200_KX_SLOT * 2 * KS, acc[2]);
    }
    int mask = 0;
#pragma unroll
    for (int j = 0; j < 3; ++j) {
        if (acc[j].active) {
            mask |= 1 << j;
            T* K = Kx + B200_KX_SLOT * j * KS;
#pragma unroll
            for (int r = 0; r < 3; ++r) { K[(21 + r) * KS] = acc[j].Wn[r]; K[(24 + r) * KS] = acc[j].Wf[r]; }
        }
    }
    return mask;
}

// ---- leg-parallel formulation ------------------------------------------------------------------------------------
// The two legs never couple except through the 6 base DoFs (M[left, right] = 0, also through foot contact), so one
// physics tick splits into
//   phase 1 (per leg, independent): kinematics / inertias / bias wrenches / foot contact of the leg, its rows of the
//            mass matrix, its share of the base block and base right-hand side, L^T D L elimination of its 6 DoFs
//            (leaf to root) and the matching part of the forward substitution;
//   exchange: the two legs' partial base blocks (21 values) and base right-hand sides (6 values) are ADDED;
//   phase 2: factor / solve the 6x6 base system (done redundantly by both), back-substitute the leg, integrate.
// On the device the two phases of an environment run on a PAIR of adjacent lanes and the exchange is 27
// __shfl_xor_sync(.., 1); on the host (tests, CPU port) t1_tick() runs the two legs one after the other - same code,
// and a + b is commutative, so both produce bit-identical results for a given scalar type.
template <typename T> struct LegState {   // what one lane integrates: the base (redundantly) and its own leg
    T pos[3], quat[4], vlin[3], wb[3];
    T q[6], qd[6];
};
template <typename T> struct LegParams {  // domain-randomised per env (envs/t1.py:139-167): trunk + the 6 bodies of this leg
    T mass0, com0[3];
    T mass[6], com[6][3];
    T mu, kscale, cscale;
};
template <typename T> struct LegWork {
    T Mbb[21];     // base-base block, lower-tri, order v(3), w_body(3): partial after phase 1, total after the exchange
    T rb[6];       // base right-hand side: partial / total
    T Mlb[6][6];   // leg row k x base column j  (L factors after phase 1)
    T Mll[21];     // leg-leg lower-tri (D^-1 on the diagonal, L below, after phase 1)
    T xl[6];       // leg right-hand side after the leg's part of the forward substitution
    T foot_fn;     // explicit normal-force estimate of this foot [N]
    T body_f2[3];  // |contact force|^2 on this leg's hip-yaw link, shank and foot (net_contact_force rows, envs/t1.py:553,628)
    T trunk_f[3];  // this lane's share of the contact force on the trunk (the pair adds the two shares)
    T foot_pos[3]; // foot link origin, world
    T Rf[3][3];    // foot rotation
};
B200_HD constexpr int tri(int i, int j) { return i * (i + 1) / 2 + j; }

// ---- what the forward pass publishes for the shapes (also used by the host wrapper to pre-fill the partner leg) ----------------
// Each returns the quantity its cull needs.  R, x, w, v: frame / origin (relative to the trunk origin) / spatial velocity of
// the link; R0 = trunk frame.
// hip-yaw (ci = 0) or shank (ci = 1) cylinder: height of its lowest point above the trunk origin's height
template <typename T, typename Model> B200_HD T cylinder_low(const Model& m, int ci, const T R[3][3], const T* x) {
    const T* cp = m.cyl_pos[ci];
    const T cz = x[2] + R[2][0] * cp[0] + R[2][1] * cp[1] + R[2][2] * cp[2];
    // exact: cz - half |a_z| - rad sqrt(1 - a_z^2); the cull drops the square root (conservative by at most one radius)
    return cz - (m.cyl_half[ci] * b_abs(R[2][2]) + m.cyl_radius[ci]);
}
template <int KS, typename T, typename Model>
B200_HD void publish_hipyaw(const Model& m, const T R[3][3], const T* x, const T* w, const T* v, T* G) {
    const T* cp = m.cyl_pos[0];
#pragma unroll
    for (int r = 0; r < 3; ++r) {
        G[(24 + r) * KS] = x[r] + R[r][0] * cp[0] + R[r][1] * cp[1] + R[r][2] * cp[2];
        G[(27 + r) * KS] = R[r][2];
        G[(30 + r) * KS] = w[r];
        G[(33 + r) * KS] = v[r];
    }
}
// shank capsule: returns `inward` = how far it stays on its own side of the trunk's sagittal plane; the legs can touch only
// if inward(left leg) + inward(right leg) <= 0
template <int KS, typename T, typename Model>
B200_HD T publish_shank(const Model& m, const T R[3][3], const T* x, const T* w, const T* v, const T R0[3][3], int side, T* G) {
    const T* cp = m.cyl_pos[1];
    const T h = m.cyl_half[1];
    T yc = 0, ya = 0;
#pragma unroll
    for (int r = 0; r < 3; ++r) {
        const T c = x[r] + R[r][0] * cp[0] + R[r][1] * cp[1] + R[r][2] * cp[2];
        G[r * KS] = c + h * R[r][2];
        G[(3 + r) * KS] = c - h * R[r][2];
        G[(6 + r) * KS] = w[r];
        G[(9 + r) * KS] = v[r];
        yc += c * R0[r][1];
        ya += R[r][2] * R0[r][1];
    }
    return (side == 0 ? yc : -yc) - h * b_abs(ya) - m.cyl_radius[1];
}
template <int KS, typename T, typename Model>
B200_HD T publish_foot(const Model& m, const T R[3][3], const T* x, const T* w, const T* v, const T R0[3][3], int side, T* G) {
    T y0 = 0, y1 = 0;
#pragma unroll
    for (int r = 0; r < 3; ++r) {
        const T p0 = x[r] + R[r][0] * m.foot_cap[0][0] + R[r][1] * m.foot_cap[0][1] + R[r][2] * m.foot_cap[0][2];
        const T p1 = x[r] + R[r][0] * m.foot_cap[1][0] + R[r][1] * m.foot_cap[1][1] + R[r][2] * m.foot_cap[1][2];
        G[(12 + r) * KS] = p0;
        G[(15 + r) * KS] = p1;
        G[(18 + r) * KS] = w[r];
        G[(21 + r) * KS] = v[r];
        y0 += p0 * R0[r][1];
        y1 += p1 * R0[r][1];
    }
    if (side != 0) { y0 = -y0; y1 = -y1; }
    return (y0 < y1 ? y0 : y1) - m.foot_cap_radius;
}

// Host wrapper only: publish a leg's geometry from its state (the device publishes from inside the forward pass; the host
// runs the two legs one after the other, so the partner's geometry must exist before the first leg's tick).
template <typename T, typename Model> B200_HD void leg_geom_publish(const Model& m, const LegState<T>& s, int side, T* G) {
    T quat[4];
    {
        const T n = b_sqrt(s.quat[0] * s.quat[0] + s.quat[1] * s.quat[1] + s.quat[2] * s.quat[2] + s.quat[3] * s.quat[3]);
        const T inv = T(1) / n;
#pragma unroll
        for (int i = 0; i < 4; ++i) quat[i] = s.quat[i] * inv;
    }
    T R0[3][3], R[3][3], xp[3] = {0, 0, 0}, w[3], v[3];
    quat_to_mat(quat, R0);
#pragma unroll
    for (int r = 0; r < 3; ++r) {
#pragma unroll
        for (int c = 0; c < 3; ++c) R[r][c] = R0[r][c];
        w[r] = R0[r][0] * s.wb[0] + R0[r][1] * s.wb[1] + R0[r][2] * s.wb[2];
        v[r] = s.vlin[r];
    }
    T inward = T(1e30);
#pragma unroll
    for (int k = 0; k < 6; ++k) {
        const int b = 1 + 6 * side + k;
        const int axis = (k == 0 || k == 3 || k == 4) ? 1 : (k == 2 ? 2 : 0);
        const T* off = m.body_pos[b];
        T x[3], a[3], sl[3];
#pragma unroll
        for (int r = 0; r < 3; ++r) {
            x[r] = xp[r] + R[r][0] * off[0] + R[r][1] * off[1] + R[r][2] * off[2];
            a[r] = R[r][axis];
        }
        rotate_about_axis(R, axis, s.q[k]);
        cross3(x, a, sl);
        const T qd = s.qd[k];
#pragma unroll
        for (int r = 0; r < 3; ++r) { w[r] += qd * a[r]; v[r] += qd * sl[r]; xp[r] = x[r]; }
        if (k == 2) publish_hipyaw<1>(m, R, x, w, v, G);
        if (k == 3) inward = publish_shank<1>(m, R, x, w, v, R0, side, G);
        if (k == 5) {
            const T lo = publish_foot<1>(m, R, x, w, v, R0, side, G);
            inward = lo < inward ? lo : inward;
        }
    }
    G[36] = inward;
}

// Kx: the shape scratch described above (B200_KX_SIZE entries, stride KS) of this leg lane; Kp: the partner leg lane's.
template <typename T, int KS = 1, typename Model, typename Terr>
B200_HD void t1_leg_phase1(const Model& m, const LegParams<T>& par, LegState<T>& s, int side, const T* tau /*6*/, const T* push_f,
                           const T* push_t, const Terr& terr, LegWork<T>& W, T* Kx, const T* Kp) {
    const T dt = m.dt;
    {
        const T inv = b_rsqrt(s.quat[0] * s.quat[0] + s.quat[1] * s.quat[1] + s.quat[2] * s.quat[2] + s.quat[3] * s.quat[3]);
#pragma unroll
        for (int i = 0; i < 4; ++i) s.quat[i] *= inv;
    }
    T R0[3][3];
    quat_to_mat(s.quat, R0);
    const T zero3[3] = {0, 0, 0};
    const bool own_trunk = (side == 0);  // the trunk's own inertia / bias wrench / push is counted once, by the left leg

    // --- trunk -------------------------------------------------------------------------------------------------
    SpInertia<T> Ic0;
    si_from_body(R0, zero3, par.com0, m.inertia[0], par.mass0, par.mass0 / m.mass[0], Ic0);
    T w0[3], v0[3];
#pragma unroll
    for (int r = 0; r < 3; ++r) {
        w0[r] = R0[r][0] * s.wb[0] + R0[r][1] * s.wb[1] + R0[r][2] * s.wb[2];
        v0[r] = s.vlin[r];
    }
    T al0[3] = {0, 0, 0}, av0[3];
    cross3(v0, w0, av0);  // spatial acceleration of the base with qacc = 0: [0 ; v x w] (+ gravity as fictitious acceleration)
    av0[2] += m.gravity;
    T f0n[3] = {0, 0, 0}, f0f[3] = {0, 0, 0};  // this lane's share of the bias wrench on the trunk about the reference point
    if (own_trunk) {
        T n1[3], f1[3], n2[3], f2[3], t1[3], t2[3];
        si_mul(Ic0, al0, av0, n1, f1);
        si_mul(Ic0, w0, v0, n2, f2);
        cross3(w0, n2, t1);
        cross3(v0, f2, t2);
#pragma unroll
        for (int r = 0; r < 3; ++r) f0n[r] = n1[r] + t1[r] + t2[r];
        cross3(w0, f2, t1);
#pragma unroll
        for (int r = 0; r < 3; ++r) f0f[r] = f1[r] + t1[r];
        // push on the trunk (local frame, at the CoM; envs/t1.py:522-527): subtract from the bias wrench
        T Tl[3];
        cross3(par.com0, push_f, Tl);
#pragma unroll
        for (int r = 0; r < 3; ++r) Tl[r] += push_t[r];
#pragma unroll
        for (int r = 0; r < 3; ++r) {
            f0f[r] -= R0[r][0] * push_f[0] + R0[r][1] * push_f[1] + R0[r][2] * push_f[2];
            f0n[r] -= R0[r][0] * Tl[0] + R0[r][1] * Tl[1] + R0[r][2] * Tl[2];
        }
    } else {
        Ic0.m = 0;
#pragma unroll
        for (int r = 0; r < 3; ++r) Ic0.h[r] = 0;
#pragma unroll
        for (int r = 0; r < 6; ++r) Ic0.I[r] = 0;
    }
    // --- trunk box against the ground (resources/T1/T1_locomotion.xml:42): this lane takes the 4 corners on its side ------
    bool act_trunk = false, act_cyl[2] = {false, false};
#pragma unroll
    for (int r = 0; r < 3; ++r) W.trunk_f[r] = 0;
    W.body_f2[0] = 0; W.body_f2[1] = 0; W.body_f2[2] = 0;
    // culls of the rarely touching shapes (evaluated after the forward pass): bit 0 trunk box, 1 hip-yaw cylinder, 2 shank
    // cylinder reach down to the highest terrain sample, bit 3 the legs are not separated by the trunk's sagittal plane
    int shape_what = 0;
    T inward = T(1e30);
    T* G = Kx + B200_KX_GEOM * KS;
    if (m.enable_body_contact) {
        // lowest of the 4 box corners on this lane's side: centre_z - |R0[2][0]| hx - |R0[2][2]| hz with the y offset signed
        const T* bp = m.trunk_box_pos;
        const T* bh = m.trunk_box_half;
        const T ly = bp[1] + (side == 0 ? bh[1] : -bh[1]);
        const T low = s.pos[2] + R0[2][0] * bp[0] + R0[2][1] * ly + R0[2][2] * bp[2] - (b_abs(R0[2][0]) * bh[0] + b_abs(R0[2][2]) * bh[2]);
        if (!(low > (T)terr.max_height)) shape_what |= 1;
    }

    // --- leg: forward pass (kinematics, inertias, bias wrenches) ------------------------------------------------------
    T ax[6][3], xj[6][3];  // world joint axis, joint anchor relative to the reference point (trunk origin)
    SpInertia<T> Ic[6];
    T fn[6][3], ff[6][3];  // bias wrench per body (angular, linear)
    T R[3][3];
#pragma unroll
    for (int r = 0; r < 3; ++r)
#pragma unroll
        for (int c = 0; c < 3; ++c) R[r][c] = R0[r][c];
    T xp[3] = {0, 0, 0};
    T wp[3] = {w0[0], w0[1], w0[2]}, vp[3] = {v0[0], v0[1], v0[2]};
    T alp[3] = {al0[0], al0[1], al0[2]}, avp[3] = {av0[0], av0[1], av0[2]};
#pragma unroll
    for (int k = 0; k < 6; ++k) {
        const int b = 1 + 6 * side + k;
        const int axis = (k == 0 || k == 3 || k == 4) ? 1 : (k == 2 ? 2 : 0);  // y x z y y x (t1_model.json "axis")
        const T* off = m.body_pos[b];
        T x[3], a[3], sl[3];
#pragma unroll
        for (int r = 0; r < 3; ++r) {
            x[r] = xp[r] + R[r][0] * off[0] + R[r][1] * off[1] + R[r][2] * off[2];
            a[r] = R[r][axis];
        }
        rotate_about_axis(R, axis, s.q[k]);
        cross3(x, a, sl);  // linear part of the motion axis about the reference point
        const T qd = s.qd[k];
        T t1[3], t2[3], t3[3];
        cross3(wp, a, t1);
        cross3(wp, sl, t2);
        cross3(vp, a, t3);
        T al[3], av[3], w[3], v[3];
#pragma unroll
        for (int r = 0; r < 3; ++r) {
            al[r] = alp[r] + qd * t1[r];
            av[r] = avp[r] + qd * (t2[r] + t3[r]);
            w[r] = wp[r] + qd * a[r];
            v[r] = vp[r] + qd * sl[r];
        }
        si_from_body(R, x, par.com[k], m.inertia[b], par.mass[k], par.mass[k] / m.mass[b], Ic[k]);
        T n1[3], f1[3], n2[3], f2[3];
        si_mul(Ic[k], al, av, n1, f1);
        si_mul(Ic[k], w, v, n2, f2);
        cross3(w, n2, t1);
        cross3(v, f2, t2);
        cross3(w, f2, t3);
#pragma unroll
        for (int r = 0; r < 3; ++r) {
            fn[k][r] = n1[r] + t1[r] + t2[r];
            ff[k][r] = f1[r] + t3[r];
            ax[k][r] = a[r];
            xj[k][r] = x[r];
            xp[r] = x[r]; wp[r] = w[r]; vp[r] = v[r]; alp[r] = al[r]; avp[r] = av[r];
        }
        // hip-yaw (k = 2) and shank (k = 3) collision cylinders against the ground (resources/T1/T1_locomotion.xml:66,71)
        // publish what the rarely touching shapes need (shape scratch) and evaluate their culls
        if (k == 2 && m.enable_body_contact) {
            publish_hipyaw<KS>(m, R, x, w, v, G);
            if (!(s.pos[2] + cylinder_low(m, 0, R, x) > (T)terr.max_height)) shape_what |= 2;
        }
        if (k == 3 && (m.enable_body_contact | m.enable_self_contact)) {
            inward = publish_shank<KS>(m, R, x, w, v, R0, side, G);
            if (m.enable_body_contact && !(s.pos[2] + cylinder_low(m, 1, R, x) > (T)terr.max_height)) shape_what |= 4;
        }
        if (k == 5 && m.enable_self_contact) {
            const T lo = publish_foot<KS>(m, R, x, w, v, R0, side, G);
            inward = lo < inward ? lo : inward;
        }
    }
    // --- rarely touching shapes: out-of-line evaluation when a cull failed, then fold the slots in ----------------------------
    int shape_mask = 0;
    bool legs_close = false;   // the two legs are NOT separated by the trunk's sagittal plane: only then can capsules touch
    if (m.enable_body_contact | m.enable_self_contact) {
        if (m.enable_self_contact) {
            G[36 * KS] = inward;
            B200_SYNCWARP();   // the partner lane's geometry (published above) and its `inward` are visible from here on
            legs_close = !(inward + Kp[(B200_KX_GEOM + 36) * KS] > 0);
        }
        if (shape_what) {
#pragma unroll
            for (int r = 0; r < 3; ++r) {
#pragma unroll
                for (int c = 0; c < 3; ++c) G[(37 + 3 * r + c) * KS] = R0[r][c];
                G[(46 + r) * KS] = s.pos[r];
                G[(49 + r) * KS] = w0[r];
                G[(52 + r) * KS] = v0[r];
            }
            shape_mask = shapes_eval<KS>(m, side, shape_what, terr, Kx, Kp);
        }
    }
    if (shape_mask & 1) {
        act_trunk = true;
#pragma unroll
        for (int r = 0; r < 3; ++r) {
            f0n[r] -= Kx[(21 + r) * KS];
            const T f = Kx[(24 + r) * KS];
            f0f[r] -= f;
            W.trunk_f[r] = f;
        }
    }
#pragma unroll
    for (int k = 2; k <= 3; ++k) {
        if (shape_mask & (1 << (k - 1))) {
            act_cyl[k - 2] = true;
            const T* K = Kx + B200_KX_SLOT * (k - 1) * KS;
            T f2 = 0;
#pragma unroll
            for (int r = 0; r < 3; ++r) {
                fn[k][r] -= K[(21 + r) * KS];
                const T f = K[(24 + r) * KS];
                ff[k][r] -= f;
                f2 += f * f;
            }
            W.body_f2[k - 2] = f2;
        }
    }
    // --- leg-leg contacts of the shank capsule ------------------------------------------------------------------------------
    // The same-type pairs (shank - shank, foot - foot) are evaluated on EVERY tick, branch-free and back to back: the kernel's
    // time is its slowest warp's, some warp of a 4096-env rollout always holds a pair of touching legs, and a rarely taken
    // block is fetched cold from L2 every time it runs (measured per 10-tick launch, standing / falling robots: behind the
    // sagittal-plane cull 119 / 211 us, always-on 135 / 148 us; without leg-leg contacts 116 / 124 us).
    // The cross pair (shank - partner's foot) needs crossed legs: behind the sagittal-plane and bounding-sphere culls.
    bool foot_hit = false;
    T foot_pm[3], foot_n[3], foot_fmag = 0;
    if (m.enable_self_contact) {
        const T* Gp = Kp + B200_KX_GEOM * KS;
        T pm[3], n[3], fmag;
        // both same-type pairs back to back (independent chains); the foot's result is consumed in the foot section
        bool hit = capsule_pair<KS>(m, side, 0, 0, G, Gp, pm, n, fmag);
        foot_hit = capsule_pair<KS>(m, side, 1, 1, G, Gp, foot_pm, foot_n, foot_fmag);
#pragma unroll
        for (int j = 0; j < 2; ++j) {
            if (j == 1) hit = legs_close && capsule_spheres_touch<KS>(m, 0, 1, G, Gp) && capsule_pair<KS>(m, side, 0, 1, G, Gp, pm, n, fmag);
            if (hit) {
                T* K = Kx + B200_KX_SLOT * 2 * KS;
                T Wn[3] = {0, 0, 0}, Wf[3] = {0, 0, 0};
                if (!act_cyl[1]) {
#pragma unroll
                    for (int i = 0; i < B200_KX_SLOT; ++i) K[i * KS] = 0;
                    act_cyl[1] = true;
                }
                capsule_pair_apply<KS>(m, pm, n, fmag, K, Wn, Wf);
                T f2 = 0;
#pragma unroll
                for (int r = 0; r < 3; ++r) {
                    fn[3][r] -= Wn[r];
                    ff[3][r] -= Wf[r];
                    const T f = K[(24 + r) * KS] + Wf[r];   // net contact force on the shank so far (slot entry 24..26)
                    K[(24 + r) * KS] = f;
                    f2 += f * f;
                }
                W.body_f2[1] = f2;
            }
        }
    }
    // --- foot contact: 4 sole corners against the heightfield --------------------------------------------------------
    T Kc[21];  // implicit contact matrix of this foot (6x6 symmetric, lower-tri, order [ang; lin]), already * dt
#pragma unroll
    for (int i = 0; i < 21; ++i) Kc[i] = 0;
#pragma unroll
    for (int r = 0; r < 3; ++r) {
        W.foot_pos[r] = s.pos[r] + xp[r];
#pragma unroll
        for (int c = 0; c < 3; ++c) W.Rf[r][c] = R[r][c];
    }
    bool foot_active = false;
    W.foot_fn = 0;
    if (m.enable_contact) {
        T Wn[3] = {0, 0, 0}, Wf[3] = {0, 0, 0};
        const T kn = m.contact_k * par.kscale, cn = m.contact_c * par.cscale;
        const T dn = cn + dt * kn;
#pragma unroll
        for (int c = 0; c < 4; ++c) {
            const T* rc = m.foot_corner[c];
            T p[3];
#pragma unroll
            for (int r = 0; r < 3; ++r) p[r] = xp[r] + R[r][0] * rc[0] + R[r][1] * rc[1] + R[r][2] * rc[2];
            const T ground = (T)terr((float)(s.pos[0] + p[0]), (float)(s.pos[1] + p[1]));
            const T depth = ground - (s.pos[2] + p[2]);
            if (depth > 0 && contact_ground_point<1>(p, depth, wp, vp, kn, cn, dn, par.mu, dt, m.stiction_vel, Kc, Wn, Wf, W.foot_fn))
                foot_active = true;
        }
        if (m.enable_self_contact) {   // foot - foot capsule pair on every tick, foot - partner's shank behind the culls (see above)
            const T* Gp = Kp + B200_KX_GEOM * KS;
            T pm[3], n[3], fmag;
            if (foot_hit) {
                capsule_pair_apply<1>(m, foot_pm, foot_n, foot_fmag, Kc, Wn, Wf);
                foot_active = true;
            }
            if (legs_close && capsule_spheres_touch<KS>(m, 1, 0, G, Gp) && capsule_pair<KS>(m, side, 1, 0, G, Gp, pm, n, fmag)) {
                capsule_pair_apply<1>(m, pm, n, fmag, Kc, Wn, Wf);
                foot_active = true;
            }
        }
        W.body_f2[2] = Wf[0] * Wf[0] + Wf[1] * Wf[1] + Wf[2] * Wf[2];
        if (foot_active) {
#pragma unroll
            for (int r = 0; r < 3; ++r) { fn[5][r] -= Wn[r]; ff[5][r] -= Wf[r]; }
        }
    }

    // --- backward pass: composite inertias, accumulated wrenches, leg right-hand side ------------------------------
#pragma unroll
    for (int k = 5; k >= 0; --k) {
        T sl[3];
        cross3(xj[k], ax[k], sl);
        W.xl[k] = tau[k] - (dot3(ax[k], fn[k]) + dot3(sl, ff[k]));
        if (k > 0) {
            si_add(Ic[k - 1], Ic[k]);
#pragma unroll
            for (int r = 0; r < 3; ++r) { fn[k - 1][r] += fn[k][r]; ff[k - 1][r] += ff[k][r]; }
        } else {
            si_add(Ic0, Ic[0]);
#pragma unroll
            for (int r = 0; r < 3; ++r) { f0n[r] += fn[0][r]; f0f[r] += ff[0][r]; }
        }
    }
#pragma unroll
    for (int r = 0; r < 3; ++r) {
        W.rb[r] = -f0f[r];
        W.rb[3 + r] = -(R0[0][r] * f0n[0] + R0[1][r] * f0n[1] + R0[2][r] * f0n[2]);
    }

    // --- mass matrix (CRBA) + implicit contact term ------------------------------------------------------------
    // Kc must be the COMPOSITE contact matrix of the subtree a joint moves: the foot's for the ankle joints, + the shank
    // cylinder's from the knee up, + the hip-yaw cylinder's from the hip-yaw joint up, + this lane's share of the trunk box for
    // the base block: the sweep runs leaf to root and adds the parked matrices (shape scratch) as it passes their bodies.
    // (Skipped cold blocks are kept SHORT on purpose: the tick streams its code from L2, and a taken branch over more than a few
    // instruction lines costs a fetch round trip - measured 4 % of the tick for four 100-instruction blocks.)
    bool kact = foot_active;
#pragma unroll
    for (int k = 5; k >= 0; --k) {
        if ((k == 3 || k == 2) && act_cyl[k - 2]) {
#pragma unroll
            for (int i = 0; i < 21; ++i) Kc[i] += Kx[(B200_KX_SLOT * (k - 1) + i) * KS];
            kact = true;
        }
        T sl[3], Fn[3], Ff[3];
        cross3(xj[k], ax[k], sl);
        si_mul(Ic[k], ax[k], sl, Fn, Ff);
        if (kact) {
            const T S6[6] = {ax[k][0], ax[k][1], ax[k][2], sl[0], sl[1], sl[2]};
#pragma unroll
            for (int r = 0; r < 6; ++r) {
                T acc = 0;
#pragma unroll
                for (int c = 0; c < 6; ++c) {
                    const int lo = (r < c) ? r : c, hi = (r < c) ? c : r;
                    acc += Kc[tri(hi, lo)] * S6[c];
                }
                if (r < 3) Fn[r] += acc; else Ff[r - 3] += acc;
            }
        }
#pragma unroll
        for (int j = 0; j < 3; ++j) {
            W.Mlb[k][j] = Ff[j];
            W.Mlb[k][3 + j] = R0[0][j] * Fn[0] + R0[1][j] * Fn[1] + R0[2][j] * Fn[2];
        }
#pragma unroll
        for (int j = 0; j <= k; ++j) {
            T slj[3];
            cross3(xj[j], ax[j], slj);
            W.Mll[tri(k, j)] = dot3(ax[j], Fn) + dot3(slj, Ff);
        }
    }
    if (act_trunk) {
#pragma unroll
        for (int i = 0; i < 21; ++i) Kc[i] += Kx[i * KS];
        kact = true;
    }
    {
        // base block share: S_lin,k = [0; e_k], S_ang,k = [R0[:,k]; 0] through (trunk share +) this leg's composite inertia (+ K)
        T Fn[6][3], Ff[6][3];
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            T e[3] = {0, 0, 0};
            e[k] = 1;
            si_mul(Ic0, zero3, e, Fn[k], Ff[k]);
            const T a[3] = {R0[0][k], R0[1][k], R0[2][k]};
            si_mul(Ic0, a, zero3, Fn[3 + k], Ff[3 + k]);
            if (kact) {
#pragma unroll
                for (int r = 0; r < 3; ++r) {
                    const int ci = 3 + k;
                    Fn[k][r] += Kc[tri(ci, r)];
                    const int lo = (r < k) ? r : k, hi = (r < k) ? k : r;
                    Ff[k][r] += Kc[tri(3 + hi, 3 + lo)];
                    T accn = 0, accf = 0;
#pragma unroll
                    for (int c = 0; c < 3; ++c) {
                        const int l2 = (r < c) ? r : c, h2 = (r < c) ? c : r;
                        accn += Kc[tri(h2, l2)] * a[c];
                        accf += Kc[tri(3 + r, c)] * a[c];
                    }
                    Fn[3 + k][r] += accn;
                    Ff[3 + k][r] += accf;
                }
            }
        }
#pragma unroll
        for (int i = 0; i < 6; ++i) {
#pragma unroll
            for (int j = 0; j <= i; ++j) {
                if (j < 3) W.Mbb[tri(i, j)] = Ff[i][j];
                else W.Mbb[tri(i, j)] = R0[0][j - 3] * Fn[i][0] + R0[1][j - 3] * Fn[i][1] + R0[2][j - 3] * Fn[i][2];
            }
        }
    }
    // --- joint limits: spring explicit + linearly-implicit damper on the diagonal -------------------------------------
    if (m.enable_limits) {
#pragma unroll
        for (int k = 0; k < 6; ++k) {
            const int j = 6 * side + k;
            const T q = s.q[k], qd = s.qd[k];
            // both legs' limits as immediate constant-bank operands + a select: indexing the bank with the lane's `side` made
            // each compare wait for a dependent LDC (4.4 % of the tick's stall samples on these two lines)
            const T lo = (side == 0) ? m.jnt_lower[k] : m.jnt_lower[6 + k];
            const T hi = (side == 0) ? m.jnt_upper[k] : m.jnt_upper[6 + k];
            T viol = 0;
            if (q < lo) viol = lo - q;
            else if (q > hi) viol = hi - q;
            if (viol != 0) {
                const T ke = m.limit_k * m.dof_inertia[j], ce = m.limit_c * m.dof_inertia[j];
                const T de = ce + dt * ke;
                W.xl[k] += ke * viol - de * qd;
                W.Mll[tri(k, k)] += dt * de;
            }
        }
    }
    // --- L^T D L elimination of the leg DoFs, leaf to root (MuJoCo mj_factorM order: no fill-in) ------------------------
#pragma unroll
    for (int k = 5; k >= 0; --k) {
        const T inv = b_div(T(1), W.Mll[tri(k, k)]);   // kept ON the diagonal below: phase 2 multiplies instead of dividing again
        // rows i = leg DoFs below k
#pragma unroll
        for (int i = 0; i < k; ++i) {
            const T l = W.Mll[tri(k, i)] * inv;
#pragma unroll
            for (int j = 0; j < 6; ++j) W.Mlb[i][j] -= l * W.Mlb[k][j];
#pragma unroll
            for (int j = 0; j <= i; ++j) W.Mll[tri(i, j)] -= l * W.Mll[tri(k, j)];
        }
        // rows i = base DoFs
#pragma unroll
        for (int i = 0; i < 6; ++i) {
            const T l = W.Mlb[k][i] * inv;
#pragma unroll
            for (int j = 0; j <= i; ++j) W.Mbb[tri(i, j)] -= l * W.Mlb[k][j];
        }
#pragma unroll
        for (int i = 0; i < k; ++i) W.Mll[tri(k, i)] *= inv;
#pragma unroll
        for (int j = 0; j < 6; ++j) W.Mlb[k][j] *= inv;
        W.Mll[tri(k, k)] = inv;
    }
    // forward substitution, leg part: x_j -= L_ij x_i for i = leg DoFs (descending)
#pragma unroll
    for (int i = 5; i >= 0; --i) {
#pragma unroll
        for (int j = 0; j < i; ++j) W.xl[j] -= W.Mll[tri(i, j)] * W.xl[i];
#pragma unroll
        for (int j = 0; j < 6; ++j) W.rb[j] -= W.Mlb[i][j] * W.xl[i];
    }
}

// phase 2: W.Mbb / W.rb now hold the SUM over both legs.  Produces qacc (base 6 + this leg 6) and integrates.
template <typename T, typename Model>
B200_HD void t1_leg_phase2(const Model& m, LegState<T>& s, LegWork<T>& W, T* qacc_base, T* qacc_leg, bool integrate) {
    const T dt = m.dt;
    T* B = W.Mbb;
    T* x = W.rb;
    // factor the 6x6 base block
#pragma unroll
    for (int k = 5; k >= 0; --k) {
        const T inv = b_div(T(1), B[tri(k, k)]);
#pragma unroll
        for (int i = 0; i < k; ++i) {
            const T l = B[tri(k, i)] * inv;
#pragma unroll
            for (int j = 0; j <= i; ++j) B[tri(i, j)] -= l * B[tri(k, j)];
        }
#pragma unroll
        for (int i = 0; i < k; ++i) B[tri(k, i)] *= inv;
        B[tri(k, k)] = inv;   // D^-1 on the diagonal
    }
#pragma unroll
    for (int i = 5; i >= 0; --i)
#pragma unroll
        for (int j = 0; j < i; ++j) x[j] -= B[tri(i, j)] * x[i];
#pragma unroll
    for (int i = 0; i < 6; ++i) x[i] *= B[tri(i, i)];
#pragma unroll
    for (int i = 0; i < 6; ++i)
#pragma unroll
        for (int j = 0; j < i; ++j) x[i] -= B[tri(i, j)] * x[j];
    // leg: scale by D, then x_i -= sum_j L_ij x_j over base and lower leg DoFs
#pragma unroll
    for (int i = 0; i < 6; ++i) W.xl[i] *= W.Mll[tri(i, i)];
#pragma unroll
    for (int i = 0; i < 6; ++i) {
#pragma unroll
        for (int j = 0; j < 6; ++j) W.xl[i] -= W.Mlb[i][j] * x[j];
#pragma unroll
        for (int j = 0; j < i; ++j) W.xl[i] -= W.Mll[tri(i, j)] * W.xl[j];
    }
#pragma unroll
    for (int i = 0; i < 6; ++i) { qacc_base[i] = x[i]; qacc_leg[i] = W.xl[i]; }
    if (!integrate) return;
    // semi-implicit Euler (MuJoCo mj_Euler): velocities first, positions with the NEW velocities
#pragma unroll
    for (int r = 0; r < 3; ++r) {
        s.vlin[r] += dt * x[r];
        s.wb[r] += dt * x[3 + r];
        s.pos[r] += dt * s.vlin[r];
    }
#pragma unroll
    for (int j = 0; j < 6; ++j) {
        s.qd[j] += dt * W.xl[j];
        s.q[j] += dt * s.qd[j];
    }
    {
        // quat <- quat (x) exp(dt * wb / 2)   (body-frame angular velocity: right multiplication)
        const T wn2 = s.wb[0] * s.wb[0] + s.wb[1] * s.wb[1] + s.wb[2] * s.wb[2];
        const T iw = b_rsqrt(b_max(wn2, T(1e-30))), wn = wn2 * iw;
        T sh, ch, k;
        b_sincos(T(0.5) * dt * wn, sh, ch);
        k = (wn > T(1e-9)) ? sh * iw : T(0.5) * dt;
        const T dx = k * s.wb[0], dy = k * s.wb[1], dz = k * s.wb[2], dw = ch;
        const T qx = s.quat[0], qy = s.quat[1], qz = s.quat[2], qw = s.quat[3];
        T nq[4];
        nq[0] = qw * dx + qx * dw + qy * dz - qz * dy;
        nq[1] = qw * dy - qx * dz + qy * dw + qz * dx;
        nq[2] = qw * dz + qx * dy - qy * dx + qz * dw;
        nq[3] = qw * dw - qx * dx - qy * dy - qz * dz;
        const T inv = b_rsqrt(nq[0] * nq[0] + nq[1] * nq[1] + nq[2] * nq[2] + nq[3] * nq[3]);
#pragma unroll
        for (int i = 0; i < 4; ++i) s.quat[i] = nq[i] * inv;
    }
}

// ---- whole-robot wrapper (host tests, CPU port): both legs one after the other ----------------------------------------
template <typename T> struct DynState {
    T pos[3];   // trunk origin, world
    T quat[4];  // xyzw
    T vlin[3];  // world
    T wb[3];    // angular velocity, BODY frame (MuJoCo qvel[3:6])
    T q[12], qd[12];
};
template <typename T> struct DynParams {  // domain-randomised per env (envs/t1.py:139-167)
    T mass[B200_NB];
    T com[B200_NB][3];
    T mu[2];      // Coulomb coefficient foot-ground (combined)
    T kscale[2];  // contact stiffness multiplier (1 / compliance sample)
    T cscale[2];  // contact damping multiplier (from restitution sample)
};
template <typename T> struct DynAux {  // by-products of the last tick
    T qacc[B200_NV];
    T foot_fn[2];  // explicit normal-force estimate per foot [N]
    T body_f2[2][3];  // |contact force|^2 on hip-yaw link, shank, foot of each leg
    T trunk_f[3];     // contact force on the trunk
};
template <typename T> struct MLocal {  // kept for source compatibility of callers; the leg-parallel tick needs no external storage
    T unused;
};

template <typename T> B200_HD void leg_split(const DynState<T>& s, const DynParams<T>& par, int side, LegState<T>& ls, LegParams<T>& lp) {
#pragma unroll
    for (int r = 0; r < 3; ++r) { ls.pos[r] = s.pos[r]; ls.vlin[r] = s.vlin[r]; ls.wb[r] = s.wb[r]; lp.com0[r] = par.com[0][r]; }
#pragma unroll
    for (int r = 0; r < 4; ++r) ls.quat[r] = s.quat[r];
    lp.mass0 = par.mass[0];
#pragma unroll
    for (int k = 0; k < 6; ++k) {
        ls.q[k] = s.q[6 * side + k];
        ls.qd[k] = s.qd[6 * side + k];
        lp.mass[k] = par.mass[1 + 6 * side + k];
#pragma unroll
        for (int r = 0; r < 3; ++r) lp.com[k][r] = par.com[1 + 6 * side + k][r];
    }
    lp.mu = par.mu[side]; lp.kscale = par.kscale[side]; lp.cscale = par.cscale[side];
}

// One tick of the whole robot.  `tau` = 12 joint torques (already clipped), push_f/push_t = force/torque on the trunk in
// its LOCAL frame applied at the trunk CoM.  terr(xw, yw) returns the ground height.  If `integrate` is false only qacc is
// produced (one-step parity tests).
template <typename T, typename Model, typename Terr, typename MS>
B200_HD void t1_tick(const Model& m, const DynParams<T>& par, DynState<T>& s, const T* tau, const T* push_f, const T* push_t,
                     const Terr& terr, MS&, DynAux<T>& aux, bool integrate) {
    LegState<T> ls[2];
    LegParams<T> lp[2];
    LegWork<T> W[2];
    T Kx[2][B200_KX_SIZE];
    for (int side = 0; side < 2; ++side) {
        leg_split(s, par, side, ls[side], lp[side]);
        leg_geom_publish(m, ls[side], side, Kx[side] + B200_KX_GEOM);   // the device publishes from inside the forward pass
    }
    for (int side = 0; side < 2; ++side)
        t1_leg_phase1<T>(m, lp[side], ls[side], side, tau + 6 * side, push_f, push_t, terr, W[side], Kx[side], Kx[1 - side]);
    for (int i = 0; i < 21; ++i) { const T t = W[0].Mbb[i] + W[1].Mbb[i]; W[0].Mbb[i] = t; W[1].Mbb[i] = t; }
    for (int i = 0; i < 6; ++i) { const T t = W[0].rb[i] + W[1].rb[i]; W[0].rb[i] = t; W[1].rb[i] = t; }
    for (int side = 0; side < 2; ++side) {
        T qb[6], ql[6];
        t1_leg_phase2<T>(m, ls[side], W[side], qb, ql, integrate);
        for (int i = 0; i < 6; ++i) { aux.qacc[i] = qb[i]; aux.qacc[6 + 6 * side + i] = ql[i]; }
        aux.foot_fn[side] = W[side].foot_fn;
        for (int i = 0; i < 3; ++i) aux.body_f2[side][i] = W[side].body_f2[i];
        for (int k = 0; k < 6; ++k) { s.q[6 * side + k] = ls[side].q[k]; s.qd[6 * side + k] = ls[side].qd[k]; }
    }
    for (int r = 0; r < 3; ++r) { s.pos[r] = ls[0].pos[r]; s.vlin[r] = ls[0].vlin[r]; s.wb[r] = ls[0].wb[r]; }
    for (int r = 0; r < 4; ++r) s.quat[r] = ls[0].quat[r];
    for (int r = 0; r < 3; ++r) aux.trunk_f[r] = W[0].trunk_f[r] + W[1].trunk_f[r];
}

// Pose (world position + quaternion xyzw) of ONE foot link from a lane's state.
template <typename T, typename Model>
B200_HD void t1_foot_fk(const Model& m, const LegState<T>& s, int side, T* foot_pos, T* foot_quat) {
    T R[3][3];
    quat_to_mat(s.quat, R);
    T x[3] = {s.pos[0], s.pos[1], s.pos[2]};
#pragma unroll
    for (int k = 0; k < 6; ++k) {
        const int b = 1 + 6 * side + k;
        const int axis = (k == 0 || k == 3 || k == 4) ? 1 : (k == 2 ? 2 : 0);
        const T* off = m.body_pos[b];
        T nx[3];
#pragma unroll
        for (int r = 0; r < 3; ++r) nx[r] = x[r] + R[r][0] * off[0] + R[r][1] * off[1] + R[r][2] * off[2];
#pragma unroll
        for (int r = 0; r < 3; ++r) x[r] = nx[r];
        rotate_about_axis(R, axis, s.q[k]);
    }
#pragma unroll
    for (int r = 0; r < 3; ++r) foot_pos[r] = x[r];
    mat_to_quat(R, foot_quat);
}

// Feet poses (world position + quaternion xyzw) from the current configuration: the rigid_body_state rows that
// envs/t1.py:223-224,530-531 read for the two foot links.
template <typename T, typename Model>
B200_HD void t1_feet_fk(const Model& m, const DynState<T>& s, T foot_pos[2][3], T foot_quat[2][4]) {
    T R0[3][3];
    quat_to_mat(s.quat, R0);
#pragma unroll
    for (int sgn = 0; sgn < 2; ++sgn) {
        T R[3][3];
#pragma unroll
        for (int r = 0; r < 3; ++r)
#pragma unroll
            for (int c = 0; c < 3; ++c) R[r][c] = R0[r][c];
        T x[3] = {s.pos[0], s.pos[1], s.pos[2]};
#pragma unroll
        for (int k = 0; k < 6; ++k) {
            const int b = 1 + 6 * sgn + k;
            const int axis = (k == 0 || k == 3 || k == 4) ? 1 : (k == 2 ? 2 : 0);
            const T* off = m.body_pos[b];
            T nx[3];
#pragma unroll
            for (int r = 0; r < 3; ++r) nx[r] = x[r] + R[r][0] * off[0] + R[r][1] * off[1] + R[r][2] * off[2];
#pragma unroll
            for (int r = 0; r < 3; ++r) x[r] = nx[r];
            rotate_about_axis(R, axis, s.q[6 * sgn + k]);
        }
#pragma unroll
        for (int r = 0; r < 3; ++r) foot_pos[sgn][r] = x[r];
        mat_to_quat(R, foot_quat[sgn]);
    }
}

}  // namespace b200
