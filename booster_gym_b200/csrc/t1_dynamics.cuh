// t1_dynamics.cuh - one physics tick (2 ms) of the Booster T1 tree for ONE environment.
//
// Replaces what `gym.simulate` does between envs/t1.py:450-455 (PhysX, closed source) and what `mujoco.mj_step`
// does at play_mujoco.py:756, on the model of resources/T1/T1_locomotion.xml:37-119.  Smooth dynamics follow the
// MuJoCo forward pipeline: generalised velocity [v_world, omega_body, qd], world-aligned
// spatial algebra about a common reference point (here the trunk origin instead of the subtree CoM, which keeps
// fp32 numbers small on an 80 m terrain), composite-rigid-body mass matrix, recursive Newton-Euler bias, sparse
// L^T D L factorisation in leaf-to-root order (left/right leg blocks never couple), semi-implicit Euler with an
// exact quaternion exponential.  Contact is NOT MuJoCo's convex solver: 8 sole corners against the heightfield with
// a linearly-implicit spring-damper normal force and a lagged, regularised Coulomb friction (DESIGN.md "contact").
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
#else
#define B200_HD inline
#endif

namespace b200 {

template <typename T> struct ModelOf;
template <> struct ModelOf<float> { typedef B200T1ModelF type; };
template <> struct ModelOf<double> { typedef B200T1ModelD type; };

B200_HD void b_sincos(float x, float& s, float& c) {
#if defined(__CUDA_ARCH__)
    sincosf(x, &s, &c);
#else
    s = sinf(x); c = cosf(x);
#endif
}
B200_HD void b_sincos(double x, double& s, double& c) { s = sin(x); c = cos(x); }
B200_HD float b_sqrt(float x) { return sqrtf(x); }
B200_HD double b_sqrt(double x) { return sqrt(x); }
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
    T tr = R[0][0] + R[1][1] + R[2][2];
    if (tr > 0) {
        T s = b_sqrt(tr + 1) * 2;
        q[3] = s / 4; q[0] = (R[2][1] - R[1][2]) / s; q[1] = (R[0][2] - R[2][0]) / s; q[2] = (R[1][0] - R[0][1]) / s;
    } else if (R[0][0] > R[1][1] && R[0][0] > R[2][2]) {
        T s = b_sqrt(1 + R[0][0] - R[1][1] - R[2][2]) * 2;
        q[3] = (R[2][1] - R[1][2]) / s; q[0] = s / 4; q[1] = (R[0][1] + R[1][0]) / s; q[2] = (R[0][2] + R[2][0]) / s;
    } else if (R[1][1] > R[2][2]) {
        T s = b_sqrt(1 + R[1][1] - R[0][0] - R[2][2]) * 2;
        q[3] = (R[0][2] - R[2][0]) / s; q[0] = (R[0][1] + R[1][0]) / s; q[1] = s / 4; q[2] = (R[1][2] + R[2][1]) / s;
    } else {
        T s = b_sqrt(1 + R[2][2] - R[0][0] - R[1][1]) * 2;
        q[3] = (R[1][0] - R[0][1]) / s; q[0] = (R[0][2] + R[2][0]) / s; q[1] = (R[1][2] + R[2][1]) / s; q[2] = s / 4;
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

// ---- per-environment data ---------------------------------------------------------------------------------------
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
};

// mass matrix storage: plain lower-triangular array (host / local memory variant)
template <typename T> struct MLocal {
    T a[B200_NV * (B200_NV + 1) / 2];
    B200_HD T& operator()(int i, int j) { return a[i * (i + 1) / 2 + j]; }
};

B200_HD constexpr bool dof_coupled(int i, int j) {
    // base 0..5, left leg 6..11, right leg 12..17: the two legs never couple (also not through foot contact)
    return !((i >= 6 && i < 12 && j >= 12) || (j >= 6 && j < 12 && i >= 12));
}

// Sparse L^T D L in place (MuJoCo mj_factorM order: eliminate from the last DoF down, so the tree causes no fill-in).
// Afterwards M(k,k) holds D_k and M(k,i), i<k holds L_ki.
template <typename T, typename MS> B200_HD void factor_ltdl(MS& M) {
#pragma unroll
    for (int k = B200_NV - 1; k >= 0; --k) {
        T inv = T(1) / M(k, k);
#pragma unroll
        for (int i = 0; i < k; ++i) {
            if (!dof_coupled(k, i)) continue;
            T l = M(k, i) * inv;
#pragma unroll
            for (int j = 0; j <= i; ++j) {
                if (!dof_coupled(k, j) || !dof_coupled(i, j)) continue;
                M(i, j) -= l * M(k, j);
            }
        }
#pragma unroll
        for (int i = 0; i < k; ++i)
            if (dof_coupled(k, i)) M(k, i) *= inv;
    }
}
template <typename T, typename MS> B200_HD void solve_ltdl(MS& M, T* x) {
#pragma unroll
    for (int i = B200_NV - 1; i >= 0; --i) {
#pragma unroll
        for (int j = 0; j < i; ++j)
            if (dof_coupled(i, j)) x[j] -= M(i, j) * x[i];
    }
#pragma unroll
    for (int i = 0; i < B200_NV; ++i) x[i] /= M(i, i);
#pragma unroll
    for (int i = 0; i < B200_NV; ++i) {
#pragma unroll
        for (int j = 0; j < i; ++j)
            if (dof_coupled(i, j)) x[i] -= M(i, j) * x[j];
    }
}

// One tick.  `tau` = 12 joint torques (already clipped), push_f/push_t = force/torque on the trunk in its LOCAL frame
// applied at the trunk CoM (envs/t1.py:522-527, gymapi.LOCAL_SPACE).  terr(xw, yw) returns the ground height.
// If `integrate` is false only qacc is produced (used by the one-step parity tests).
template <typename T, typename Model, typename Terr, typename MS>
B200_HD void t1_tick(const Model& m, const DynParams<T>& par, DynState<T>& s, const T* tau, const T* push_f,
                     const T* push_t, const Terr& terr, MS& M, DynAux<T>& aux, bool integrate) {
    const T dt = m.dt;
    // --- normalise the base quaternion, base rotation ---------------------------------------------------------
    {
        T n = b_sqrt(s.quat[0] * s.quat[0] + s.quat[1] * s.quat[1] + s.quat[2] * s.quat[2] + s.quat[3] * s.quat[3]);
        T inv = T(1) / n;
#pragma unroll
        for (int i = 0; i < 4; ++i) s.quat[i] *= inv;
    }
    T R0[3][3];
    quat_to_mat(s.quat, R0);
    const T zero3[3] = {0, 0, 0};

    // --- trunk: spatial inertia, velocity, bias acceleration -----------------------------------------------------
    SpInertia<T> Ic0;
    si_from_body(R0, zero3, par.com[0], m.inertia[0], par.mass[0], par.mass[0] / m.mass[0], Ic0);
    T w0[3], v0[3];
#pragma unroll
    for (int r = 0; r < 3; ++r) {
        w0[r] = R0[r][0] * s.wb[0] + R0[r][1] * s.wb[1] + R0[r][2] * s.wb[2];
        v0[r] = s.vlin[r];
    }
    T al0[3] = {0, 0, 0}, av0[3];
    cross3(v0, w0, av0);  // spatial acceleration of the base with qacc = 0:  [0 ; v x w]  (+ gravity as fictitious accel)
    av0[2] += m.gravity;
    T f0n[3], f0f[3];  // accumulated bias wrench on the trunk about the reference point
    {
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
    }
    // push on the trunk (local frame, at the CoM): subtract from the bias wrench
    {
        T Fw[3], Tl[3], Tw[3], cw[3], cxF[3];
        cross3(par.com[0], push_f, Tl);
#pragma unroll
        for (int r = 0; r < 3; ++r) Tl[r] += push_t[r];
#pragma unroll
        for (int r = 0; r < 3; ++r) {
            Fw[r] = R0[r][0] * push_f[0] + R0[r][1] * push_f[1] + R0[r][2] * push_f[2];
            Tw[r] = R0[r][0] * Tl[0] + R0[r][1] * Tl[1] + R0[r][2] * Tl[2];
        }
        (void)cw; (void)cxF;
#pragma unroll
        for (int r = 0; r < 3; ++r) { f0f[r] -= Fw[r]; f0n[r] -= Tw[r]; }
    }

    // --- legs: forward pass (kinematics, inertias, bias wrenches, foot contact) -----------------------------------
    T ax[2][6][3], xj[2][6][3];  // world joint axis, joint anchor rel. reference point
    SpInertia<T> Ic[2][6];
    T fn[2][6][3], ff[2][6][3];  // bias wrench per body (angular, linear)
    T Kc[2][21];                 // implicit contact matrix per foot (6x6 sym, lower-tri, order [ang; lin]), already * dt
    T Rf[2][3][3];               // foot rotation (for the feet outputs)
    bool foot_active[2];
#pragma unroll
    for (int sgn = 0; sgn < 2; ++sgn) {
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
            const int b = 1 + 6 * sgn + k;
            const int axis = (k == 0 || k == 3 || k == 4) ? 1 : (k == 2 ? 2 : 0);  // y x z y y x (t1_model.json "axis")
            const T* off = m.body_pos[b];
            T x[3], a[3], sl[3];
#pragma unroll
            for (int r = 0; r < 3; ++r) {
                x[r] = xp[r] + R[r][0] * off[0] + R[r][1] * off[1] + R[r][2] * off[2];
                a[r] = R[r][axis];
            }
            rotate_about_axis(R, axis, s.q[6 * sgn + k]);
            cross3(x, a, sl);  // linear part of the motion axis about the reference point
            const T qd = s.qd[6 * sgn + k];
            // velocity-product acceleration: a_b = a_p + qd * (V_p x S)
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
            si_from_body(R, x, par.com[b], m.inertia[b], par.mass[b], par.mass[b] / m.mass[b], Ic[sgn][k]);
            T n1[3], f1[3], n2[3], f2[3];
            si_mul(Ic[sgn][k], al, av, n1, f1);
            si_mul(Ic[sgn][k], w, v, n2, f2);
            cross3(w, n2, t1);
            cross3(v, f2, t2);
            cross3(w, f2, t3);
#pragma unroll
            for (int r = 0; r < 3; ++r) {
                fn[sgn][k][r] = n1[r] + t1[r] + t2[r];
                ff[sgn][k][r] = f1[r] + t3[r];
                ax[sgn][k][r] = a[r];
                xj[sgn][k][r] = x[r];
                xp[r] = x[r]; wp[r] = w[r]; vp[r] = v[r]; alp[r] = al[r]; avp[r] = av[r];
            }
        }
        // ---- foot contact: 4 sole corners against the heightfield -------------------------------------------
#pragma unroll
        for (int i = 0; i < 21; ++i) Kc[sgn][i] = 0;
#pragma unroll
        for (int r = 0; r < 3; ++r)
#pragma unroll
            for (int c = 0; c < 3; ++c) Rf[sgn][r][c] = R[r][c];
        foot_active[sgn] = false;
        aux.foot_fn[sgn] = 0;
        if (m.enable_contact) {
            T Wn[3] = {0, 0, 0}, Wf[3] = {0, 0, 0};
            const T kn = m.contact_k * par.kscale[sgn], cn = m.contact_c * par.cscale[sgn];
            const T dn = cn + dt * kn;
#pragma unroll
            for (int c = 0; c < 4; ++c) {
                const T* rc = m.foot_corner[c];
                T p[3];
#pragma unroll
                for (int r = 0; r < 3; ++r) p[r] = xp[r] + R[r][0] * rc[0] + R[r][1] * rc[1] + R[r][2] * rc[2];
                const T ground = (T)terr((float)(s.pos[0] + p[0]), (float)(s.pos[1] + p[1]));
                const T depth = ground - (s.pos[2] + p[2]);
                if (depth > 0) {
                    T wxp[3];
                    cross3(wp, p, wxp);
                    const T vc[3] = {vp[0] + wxp[0], vp[1] + wxp[1], vp[2] + wxp[2]};
                    const T fn0 = kn * depth - cn * vc[2];
                    if (fn0 > 0) {
                        foot_active[sgn] = true;
                        aux.foot_fn[sgn] += fn0;
                        const T vt = b_sqrt(vc[0] * vc[0] + vc[1] * vc[1]);
                        const T dtan = par.mu[sgn] * fn0 / b_max(vt, m.stiction_vel);
                        const T Fe[3] = {-dtan * vc[0], -dtan * vc[1], kn * depth - dn * vc[2]};
                        T pxF[3];
                        cross3(p, Fe, pxF);
#pragma unroll
                        for (int r = 0; r < 3; ++r) { Wn[r] += pxF[r]; Wf[r] += Fe[r]; }
                        // K += dt * sum_d D_d r_d^T r_d with rows r_x = [0, pz, -py | 1 0 0], r_y = [-pz, 0, px | 0 1 0],
                        // r_z = [py, -px, 0 | 0 0 1]   (point velocity = v + w x p)
                        const T Dx = dt * dtan, Dy = dt * dtan, Dz = dt * dn;
                        T* K = Kc[sgn];
                        // lower-tri index (i,j) -> i(i+1)/2+j ; order: 0 wx,1 wy,2 wz,3 vx,4 vy,5 vz
                        K[0] += Dy * p[2] * p[2] + Dz * p[1] * p[1];                      // (0,0)
                        K[1] += -Dz * p[0] * p[1];                                        // (1,0)
                        K[2] += Dx * p[2] * p[2] + Dz * p[0] * p[0];                      // (1,1)
                        K[3] += -Dy * p[0] * p[2];                                        // (2,0)
                        K[4] += -Dx * p[1] * p[2];                                        // (2,1)
                        K[5] += Dx * p[1] * p[1] + Dy * p[0] * p[0];                      // (2,2)
                        K[7] += Dx * p[2];                                                // (3,1)
                        K[8] += -Dx * p[1];                                               // (3,2)
                        K[9] += Dx;                                                       // (3,3)
                        K[10] += -Dy * p[2];                                              // (4,0)
                        K[12] += Dy * p[0];                                               // (4,2)
                        K[14] += Dy;                                                      // (4,4)
                        K[15] += Dz * p[1];                                               // (5,0)
                        K[16] += -Dz * p[0];                                              // (5,1)
                        K[20] += Dz;                                                      // (5,5)
                    }
                }
            }
            if (foot_active[sgn]) {
#pragma unroll
                for (int r = 0; r < 3; ++r) { fn[sgn][5][r] -= Wn[r]; ff[sgn][5][r] -= Wf[r]; }
            }
        }
    }

    // --- backward pass: composite inertias and accumulated wrenches, bias forces ----------------------------------
    T rhs[B200_NV];
#pragma unroll
    for (int sgn = 0; sgn < 2; ++sgn) {
#pragma unroll
        for (int k = 5; k >= 0; --k) {
            const int d = 6 + 6 * sgn + k;
            T sl[3];
            cross3(xj[sgn][k], ax[sgn][k], sl);
            rhs[d] = tau[6 * sgn + k] - (dot3(ax[sgn][k], fn[sgn][k]) + dot3(sl, ff[sgn][k]));
            if (k > 0) {
                si_add(Ic[sgn][k - 1], Ic[sgn][k]);
#pragma unroll
                for (int r = 0; r < 3; ++r) { fn[sgn][k - 1][r] += fn[sgn][k][r]; ff[sgn][k - 1][r] += ff[sgn][k][r]; }
            } else {
                si_add(Ic0, Ic[sgn][0]);
#pragma unroll
                for (int r = 0; r < 3; ++r) { f0n[r] += fn[sgn][0][r]; f0f[r] += ff[sgn][0][r]; }
            }
        }
    }
#pragma unroll
    for (int r = 0; r < 3; ++r) {
        rhs[r] = -f0f[r];
        rhs[3 + r] = -(R0[0][r] * f0n[0] + R0[1][r] * f0n[1] + R0[2][r] * f0n[2]);
    }

    // --- mass matrix (CRBA) + implicit contact term ------------------------------------------------------------
    // base block: S_lin,k = [0; e_k], S_ang,k = [R0[:,k]; 0]
    {
        // F for the 6 base DoFs through the total composite inertia (+ both feet's K)
        T Fn[6][3], Ff[6][3];
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            T e[3] = {0, 0, 0};
            e[k] = 1;
            si_mul(Ic0, zero3, e, Fn[k], Ff[k]);
            const T a[3] = {R0[0][k], R0[1][k], R0[2][k]};
            si_mul(Ic0, a, zero3, Fn[3 + k], Ff[3 + k]);
#pragma unroll
            for (int sgn = 0; sgn < 2; ++sgn) {
                if (!foot_active[sgn]) continue;
                const T* K = Kc[sgn];
                // K * [0; e_k]  -> column 3+k of K ; K * [a; 0]
#pragma unroll
                for (int r = 0; r < 3; ++r) {
                    const int ci = 3 + k;
                    // symmetric fetch K(r, ci) with ci > r
                    Fn[k][r] += K[ci * (ci + 1) / 2 + r];
                    const int lo = (r < k) ? r : k, hi = (r < k) ? k : r;
                    Ff[k][r] += K[(3 + hi) * (3 + hi + 1) / 2 + 3 + lo];
                    T accn = 0, accf = 0;
#pragma unroll
                    for (int c = 0; c < 3; ++c) {
                        const int l2 = (r < c) ? r : c, h2 = (r < c) ? c : r;
                        accn += K[h2 * (h2 + 1) / 2 + l2] * a[c];
                        accf += K[(3 + r) * (3 + r + 1) / 2 + c] * a[c];
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
                // M(i,j) = S_j . F_i
                if (j < 3) M(i, j) = Ff[i][j];
                else M(i, j) = R0[0][j - 3] * Fn[i][0] + R0[1][j - 3] * Fn[i][1] + R0[2][j - 3] * Fn[i][2];
            }
        }
    }
#pragma unroll
    for (int sgn = 0; sgn < 2; ++sgn) {
#pragma unroll
        for (int k = 0; k < 6; ++k) {
            const int d = 6 + 6 * sgn + k;
            T sl[3], Fn[3], Ff[3];
            cross3(xj[sgn][k], ax[sgn][k], sl);
            si_mul(Ic[sgn][k], ax[sgn][k], sl, Fn, Ff);
            if (foot_active[sgn]) {
                const T* K = Kc[sgn];
                const T S6[6] = {ax[sgn][k][0], ax[sgn][k][1], ax[sgn][k][2], sl[0], sl[1], sl[2]};
#pragma unroll
                for (int r = 0; r < 6; ++r) {
                    T acc = 0;
#pragma unroll
                    for (int c = 0; c < 6; ++c) {
                        const int lo = (r < c) ? r : c, hi = (r < c) ? c : r;
                        acc += K[hi * (hi + 1) / 2 + lo] * S6[c];
                    }
                    if (r < 3) Fn[r] += acc; else Ff[r - 3] += acc;
                }
            }
            // with the base
#pragma unroll
            for (int j = 0; j < 3; ++j) {
                M(d, j) = Ff[j];
                M(d, 3 + j) = R0[0][j] * Fn[0] + R0[1][j] * Fn[1] + R0[2][j] * Fn[2];
            }
            // with the ancestors in the same leg (and itself)
#pragma unroll
            for (int j = 0; j <= k; ++j) {
                T slj[3];
                cross3(xj[sgn][j], ax[sgn][j], slj);
                M(d, 6 + 6 * sgn + j) = dot3(ax[sgn][j], Fn) + dot3(slj, Ff);
            }
        }
    }
    // the left/right block is structurally zero and never touched by factor/solve

    // --- joint limits: spring explicit + linearly-implicit damper on the diagonal ---------------------------------
    if (m.enable_limits) {
#pragma unroll
        for (int j = 0; j < 12; ++j) {
            const T q = s.q[j], qd = s.qd[j];
            T viol = 0;
            if (q < m.jnt_lower[j]) viol = m.jnt_lower[j] - q;
            else if (q > m.jnt_upper[j]) viol = m.jnt_upper[j] - q;
            if (viol != 0) {
                const T ke = m.limit_k * m.dof_inertia[j], ce = m.limit_c * m.dof_inertia[j];
                const T de = ce + dt * ke;
                rhs[6 + j] += ke * viol - de * qd;
                M(6 + j, 6 + j) += dt * de;
            }
        }
    }

    // --- solve, integrate (semi-implicit Euler, MuJoCo mj_Euler) --------------------------------------------------
    factor_ltdl<T>(M);
    solve_ltdl<T>(M, rhs);
#pragma unroll
    for (int i = 0; i < B200_NV; ++i) aux.qacc[i] = rhs[i];
    if (!integrate) return;
#pragma unroll
    for (int r = 0; r < 3; ++r) {
        s.vlin[r] += dt * rhs[r];
        s.wb[r] += dt * rhs[3 + r];
        s.pos[r] += dt * s.vlin[r];
    }
#pragma unroll
    for (int j = 0; j < 12; ++j) {
        s.qd[j] += dt * rhs[6 + j];
        s.q[j] += dt * s.qd[j];
    }
    {
        // quat <- quat (x) exp(dt * wb / 2)   (body-frame angular velocity: right multiplication)
        const T wn = b_sqrt(s.wb[0] * s.wb[0] + s.wb[1] * s.wb[1] + s.wb[2] * s.wb[2]);
        T sh, ch, k;
        b_sincos(T(0.5) * dt * wn, sh, ch);
        k = (wn > T(1e-9)) ? sh / wn : T(0.5) * dt;
        const T dx = k * s.wb[0], dy = k * s.wb[1], dz = k * s.wb[2], dw = ch;
        const T x = s.quat[0], y = s.quat[1], z = s.quat[2], w = s.quat[3];
        T nq[4];
        nq[0] = w * dx + x * dw + y * dz - z * dy;
        nq[1] = w * dy - x * dz + y * dw + z * dx;
        nq[2] = w * dz + x * dy - y * dx + z * dw;
        nq[3] = w * dw - x * dx - y * dy - z * dz;
        const T inv = T(1) / b_sqrt(nq[0] * nq[0] + nq[1] * nq[1] + nq[2] * nq[2] + nq[3] * nq[3]);
#pragma unroll
        for (int i = 0; i < 4; ++i) s.quat[i] = nq[i] * inv;
    }
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
