/*
 * physics_oracle.c - FP64 CPU oracle for one physics tick of the Booster T1 model.   TEST INFRASTRUCTURE ONLY.
 *
 * PARITY UNPINNED: the reference delegates this arithmetic to un-vendored third-party binaries (Isaac Gym
 * Preview 4 / PhysX at envs/t1.py:451, and the `mujoco` PyPI package - no version pin, requirements.txt:4 - at
 * play_mujoco.py:756).  Neither is installable here, and the reference has no golden trajectories.  This file
 * restates the published MuJoCo smooth-dynamics pipeline for the model constants of
 * resources/T1/T1_locomotion.xml:37-135 with a DIFFERENT ALGORITHM from the CUDA kernel:
 *
 *   kernel (csrc/t1_dynamics.cuh): spatial algebra about a common point, composite-rigid-body M, recursive
 *                                  Newton-Euler on spatial wrenches, sparse L^T D L.
 *   oracle (this file)           : classical kinematics per body, dense per-body Jacobians,
 *                                  M = sum_b Jv^T m Jv + Jw^T I Jw, bias = sum_b J^T (Newton-Euler of body b
 *                                  with qacc = 0), dense Cholesky.
 *
 * so agreement of the two on qacc checks the algebra, and the invariants in tests/ (free fall, momentum, energy)
 * check both against physics.  The contact / joint-limit force LAW (DESIGN.md "contact") is this build's own
 * and is shared as a formula, not as code.  Control law, state conventions and time step follow
 * play_mujoco.py:726-730,751-756,824 and envs/t1.py:444-456.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may call this.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include "../include/b200_t1.h"

#define NB B200_NB
#define NV B200_NV

typedef struct T1OEnv {
    double pos[3], quat[4] /*xyzw*/, vlin[3] /*world*/, wb[3] /*body*/, q[12], qd[12];
    double mass[NB], com[NB][3];
    double mu[2], kscale[2], cscale[2];
} T1OEnv;

typedef struct T1OTerrain {
    const int16_t* hf; /* NULL = plane */
    int rows, cols, border_pixels;
    float horizontal_scale;
    double vertical_scale;
} T1OTerrain;

static void cross(const double* a, const double* b, double* o) {
    double x = a[1] * b[2] - a[2] * b[1], y = a[2] * b[0] - a[0] * b[2], z = a[0] * b[1] - a[1] * b[0];
    o[0] = x; o[1] = y; o[2] = z;
}
static void matvec(const double R[9], const double* v, double* o) {
    double t[3];
    for (int r = 0; r < 3; ++r) t[r] = R[3 * r] * v[0] + R[3 * r + 1] * v[1] + R[3 * r + 2] * v[2];
    o[0] = t[0]; o[1] = t[1]; o[2] = t[2];
}
static void matmul(const double A[9], const double B[9], double C[9]) {
    double t[9];
    for (int r = 0; r < 3; ++r)
        for (int c = 0; c < 3; ++c) t[3 * r + c] = A[3 * r] * B[c] + A[3 * r + 1] * B[3 + c] + A[3 * r + 2] * B[6 + c];
    memcpy(C, t, sizeof t);
}
static void quat2mat(const double* q, double R[9]) {
    double x = q[0], y = q[1], z = q[2], w = q[3];
    R[0] = 1 - 2 * (y * y + z * z); R[1] = 2 * (x * y - w * z); R[2] = 2 * (x * z + w * y);
    R[3] = 2 * (x * y + w * z); R[4] = 1 - 2 * (x * x + z * z); R[5] = 2 * (y * z - w * x);
    R[6] = 2 * (x * z - w * y); R[7] = 2 * (y * z + w * x); R[8] = 1 - 2 * (x * x + y * y);
}
static void axis_rot(int axis, double ang, double R[9]) {
    double c = cos(ang), s = sin(ang);
    memset(R, 0, 9 * sizeof(double));
    if (axis == 0) { R[0] = 1; R[4] = c; R[5] = -s; R[7] = s; R[8] = c; }
    else if (axis == 1) { R[0] = c; R[2] = s; R[4] = 1; R[6] = -s; R[8] = c; }
    else { R[0] = c; R[1] = -s; R[3] = s; R[4] = c; R[8] = 1; }
}

/* utils/terrain.py:101-121 - fp32 index coordinate, floor, fp64 bilinear weights, int16 samples, fp32 result */
double t1o_terrain_height(const T1OTerrain* t, float px, float py) {
    if (!t || !t->hf) return 0.0;
    float x = (float)t->border_pixels + px / t->horizontal_scale;
    float y = (float)t->border_pixels + py / t->horizontal_scale;
    long x1 = (long)floorf(x), y1 = (long)floorf(y);
    long x2 = x1 + 1, y2 = y1 + 1;
    /* numpy negative-index wrap (no bounds clamp in the reference) */
    long ix1 = x1 < 0 ? x1 + t->rows : x1, ix2 = x2 < 0 ? x2 + t->rows : x2;
    long iy1 = y1 < 0 ? y1 + t->cols : y1, iy2 = y2 < 0 ? y2 + t->cols : y2;
    if (ix1 < 0 || ix2 >= t->rows || iy1 < 0 || iy2 >= t->cols || ix1 >= t->rows || iy1 >= t->cols || ix2 < 0 || iy2 < 0)
        return 0.0; /* reference would raise IndexError; callers keep robots on the field */
    double dx2 = (double)x2 - (double)x, dx1 = (double)x - (double)x1;
    double dy2 = (double)y2 - (double)y, dy1 = (double)y - (double)y1;
    double h = dx2 * dy2 * t->hf[ix1 * t->cols + iy1] + dx1 * dy2 * t->hf[ix2 * t->cols + iy1] +
               dx2 * dy1 * t->hf[ix1 * t->cols + iy2] + dx1 * dy1 * t->hf[ix2 * t->cols + iy2];
    return (double)(float)(h * t->vertical_scale);
}

static int body_parent(int b) { return (b == 0) ? -1 : ((b == 1 || b == 7) ? 0 : b - 1); }

typedef struct Kin {
    double x[NB][3], R[NB][9], a[NB][3] /*world joint axis*/, c[NB][3] /*world CoM*/, Iw[NB][9];
    double w[NB][3], al[NB][3], xdd[NB][3];
} Kin;

static void kinematics(const B200T1ModelD* m, const T1OEnv* e, Kin* k) {
    double qn = sqrt(e->quat[0] * e->quat[0] + e->quat[1] * e->quat[1] + e->quat[2] * e->quat[2] + e->quat[3] * e->quat[3]);
    double q[4] = {e->quat[0] / qn, e->quat[1] / qn, e->quat[2] / qn, e->quat[3] / qn};
    quat2mat(q, k->R[0]);
    memcpy(k->x[0], e->pos, sizeof(double) * 3);
    matvec(k->R[0], e->wb, k->w[0]);
    memset(k->al[0], 0, 24);
    memset(k->xdd[0], 0, 24);
    memset(k->a[0], 0, 24);
    for (int b = 1; b < NB; ++b) {
        int p = body_parent(b), ax = m->axis[b];
        double off[3], Rj[9], d[3], t[3], t2[3];
        matvec(k->R[p], m->body_pos[b], off);
        for (int r = 0; r < 3; ++r) k->x[b][r] = k->x[p][r] + off[r];
        for (int r = 0; r < 3; ++r) k->a[b][r] = k->R[p][3 * r + ax];
        axis_rot(ax, e->q[b - 1], Rj);
        matmul(k->R[p], Rj, k->R[b]);
        double qd = e->qd[b - 1];
        for (int r = 0; r < 3; ++r) k->w[b][r] = k->w[p][r] + k->a[b][r] * qd;
        cross(k->w[p], k->a[b], t);
        for (int r = 0; r < 3; ++r) k->al[b][r] = k->al[p][r] + t[r] * qd;
        for (int r = 0; r < 3; ++r) d[r] = k->x[b][r] - k->x[p][r];
        cross(k->al[p], d, t);
        cross(k->w[p], d, t2);
        cross(k->w[p], t2, t2);
        for (int r = 0; r < 3; ++r) k->xdd[b][r] = k->xdd[p][r] + t[r] + t2[r];
    }
    for (int b = 0; b < NB; ++b) {
        double rc[3];
        matvec(k->R[b], e->com[b], rc);
        for (int r = 0; r < 3; ++r) k->c[b][r] = k->x[b][r] + rc[r];
        const double* I = m->inertia[b];
        double sc = e->mass[b] / m->mass[b];
        double Ib[9] = {I[0] * sc, I[3] * sc, I[4] * sc, I[3] * sc, I[1] * sc, I[5] * sc, I[4] * sc, I[5] * sc, I[2] * sc};
        double Rt[9], tmp[9];
        for (int r = 0; r < 3; ++r)
            for (int c = 0; c < 3; ++c) Rt[3 * r + c] = k->R[b][3 * c + r];
        matmul(k->R[b], Ib, tmp);
        matmul(tmp, Rt, k->Iw[b]);
    }
}

/* Jacobian (3 x NV each) of a point `p` (world) rigidly attached to body `b`: linear Jv and angular Jw */
static void jacobian(const Kin* k, int b, const double* p, double Jv[3][NV], double Jw[3][NV]) {
    memset(Jv, 0, sizeof(double) * 3 * NV);
    memset(Jw, 0, sizeof(double) * 3 * NV);
    for (int i = 0; i < 3; ++i) Jv[i][i] = 1.0;
    for (int kk = 0; kk < 3; ++kk) {
        double a[3] = {k->R[0][kk], k->R[0][3 + kk], k->R[0][6 + kk]}, d[3], t[3];
        for (int r = 0; r < 3; ++r) d[r] = p[r] - k->x[0][r];
        cross(a, d, t);
        for (int r = 0; r < 3; ++r) { Jw[r][3 + kk] = a[r]; Jv[r][3 + kk] = t[r]; }
    }
    for (int j = b; j > 0; j = body_parent(j)) {
        double d[3], t[3];
        for (int r = 0; r < 3; ++r) d[r] = p[r] - k->x[j][r];
        cross(k->a[j], d, t);
        for (int r = 0; r < 3; ++r) { Jw[r][5 + j] = k->a[j][r]; Jv[r][5 + j] = t[r]; }
    }
}

static int cholesky_solve(double A[NV][NV], double* b) {
    for (int j = 0; j < NV; ++j) {
        double s = A[j][j];
        for (int k = 0; k < j; ++k) s -= A[j][k] * A[j][k];
        if (s <= 0) return -1;
        A[j][j] = sqrt(s);
        for (int i = j + 1; i < NV; ++i) {
            double t = A[i][j];
            for (int k = 0; k < j; ++k) t -= A[i][k] * A[j][k];
            A[i][j] = t / A[j][j];
        }
    }
    for (int i = 0; i < NV; ++i) {
        double t = b[i];
        for (int k = 0; k < i; ++k) t -= A[i][k] * b[k];
        b[i] = t / A[i][i];
    }
    for (int i = NV - 1; i >= 0; --i) {
        double t = b[i];
        for (int k = i + 1; k < NV; ++k) t -= A[k][i] * b[k];
        b[i] = t / A[i][i];
    }
    return 0;
}

/* mass matrix only (for the third-opinion test against tools/extract_model.py's NumPy version) */
void t1o_mass_matrix(const B200T1ModelD* m, const T1OEnv* e, double* M_out /*NV*NV*/) {
    Kin k;
    kinematics(m, e, &k);
    double M[NV][NV];
    memset(M, 0, sizeof M);
    for (int b = 0; b < NB; ++b) {
        double Jv[3][NV], Jw[3][NV];
        jacobian(&k, b, k.c[b], Jv, Jw);
        for (int i = 0; i < NV; ++i)
            for (int j = 0; j < NV; ++j) {
                double s = 0;
                for (int r = 0; r < 3; ++r) {
                    s += e->mass[b] * Jv[r][i] * Jv[r][j];
                    for (int c = 0; c < 3; ++c) s += Jw[r][i] * k.Iw[b][3 * r + c] * Jw[c][j];
                }
                M[i][j] += s;
            }
    }
    memcpy(M_out, M, sizeof M);
}

/* One penalty contact point against the ground (normal +z), the force law of DESIGN.md "contact": adds dt J^T D J to M,
 * J^T F to rhs and F to the body's net contact force.  p = world position of the point on body b. */
static double ground_point(const B200T1ModelD* m, const Kin* k, int b, const double* p, const T1OTerrain* terr, const double* qvel,
                           double kn, double cn, double mu, double M[NV][NV], double* rhs, double* body_f /*NB*3*/) {
    const double dt = m->dt, dn = cn + dt * kn;
    double Jv[3][NV], Jw[3][NV], vc[3] = {0, 0, 0};
    /* the kernel evaluates the ground at fp32(pos + rel); mirror the rounding of the lookup argument */
    double ground = t1o_terrain_height(terr, (float)p[0], (float)p[1]);
    double depth = ground - p[2];
    if (depth <= 0) return 0;
    jacobian(k, b, p, Jv, Jw);
    for (int r = 0; r < 3; ++r)
        for (int i = 0; i < NV; ++i) vc[r] += Jv[r][i] * qvel[i];
    double fn0 = kn * depth - cn * vc[2];
    if (fn0 <= 0) return 0;
    double vt = sqrt(vc[0] * vc[0] + vc[1] * vc[1]);
    double dtan = mu * fn0 / fmax(vt, m->stiction_vel);
    double D[3] = {dtan, dtan, dn};
    double Fe[3] = {-dtan * vc[0], -dtan * vc[1], kn * depth - dn * vc[2]};
    for (int i = 0; i < NV; ++i) {
        for (int r = 0; r < 3; ++r) rhs[i] += Jv[r][i] * Fe[r];
        for (int j = 0; j < NV; ++j)
            for (int r = 0; r < 3; ++r) M[i][j] += dt * D[r] * Jv[r][i] * Jv[r][j];
    }
    for (int r = 0; r < 3; ++r) body_f[3 * b + r] += Fe[r];
    return fn0;
}

/* One tick. tau[12], push_f/push_t in the trunk's LOCAL frame at the trunk CoM. Returns 0, or -1 if M_hat is not SPD.
 * qacc_out[18] (nullable); foot_fn_out[2] (nullable). */
int t1o_tick_f(const B200T1ModelD* m, T1OEnv* e, const double* tau, const double* push_f, const double* push_t,
               const T1OTerrain* terr, double* qacc_out, double* foot_fn_out, double* body_f_out /*13*3, nullable*/, int integrate) {
    Kin k;
    kinematics(m, e, &k);
    const double dt = m->dt;
    double M[NV][NV], rhs[NV], qvel[NV];
    memset(M, 0, sizeof M);
    memset(rhs, 0, sizeof rhs);
    for (int r = 0; r < 3; ++r) { qvel[r] = e->vlin[r]; qvel[3 + r] = e->wb[r]; }
    for (int j = 0; j < 12; ++j) { qvel[6 + j] = e->qd[j]; rhs[6 + j] = tau[j]; }

    for (int b = 0; b < NB; ++b) {
        double Jv[3][NV], Jw[3][NV];
        jacobian(&k, b, k.c[b], Jv, Jw);
        for (int i = 0; i < NV; ++i)
            for (int j = 0; j < NV; ++j) {
                double s = 0;
                for (int r = 0; r < 3; ++r) {
                    s += e->mass[b] * Jv[r][i] * Jv[r][j];
                    for (int c = 0; c < 3; ++c) s += Jw[r][i] * k.Iw[b][3 * r + c] * Jw[c][j];
                }
                M[i][j] += s;
            }
        /* Newton-Euler of body b with qacc = 0 */
        double rc[3], t[3], t2[3], ac[3], F[3], Iw_w[3], Ial[3], N[3];
        for (int r = 0; r < 3; ++r) rc[r] = k.c[b][r] - k.x[b][r];
        cross(k.al[b], rc, t);
        cross(k.w[b], rc, t2);
        cross(k.w[b], t2, t2);
        for (int r = 0; r < 3; ++r) ac[r] = k.xdd[b][r] + t[r] + t2[r];
        ac[2] += m->gravity;
        for (int r = 0; r < 3; ++r) F[r] = e->mass[b] * ac[r];
        matvec(k.Iw[b], k.w[b], Iw_w);
        matvec(k.Iw[b], k.al[b], Ial);
        cross(k.w[b], Iw_w, t);
        for (int r = 0; r < 3; ++r) N[r] = Ial[r] + t[r];
        for (int i = 0; i < NV; ++i)
            for (int r = 0; r < 3; ++r) rhs[i] -= Jv[r][i] * F[r] + Jw[r][i] * N[r];
    }
    /* push on the trunk at its CoM (envs/t1.py:522-527) */
    {
        double Jv[3][NV], Jw[3][NV], Fw[3], Tw[3];
        jacobian(&k, 0, k.c[0], Jv, Jw);
        matvec(k.R[0], push_f, Fw);
        matvec(k.R[0], push_t, Tw);
        for (int i = 0; i < NV; ++i)
            for (int r = 0; r < 3; ++r) rhs[i] += Jv[r][i] * Fw[r] + Jw[r][i] * Tw[r];
    }
    /* foot contact: linearly-implicit normal spring-damper + lagged regularised Coulomb friction */
    double foot_fn[2] = {0, 0};
    double body_f[NB * 3];
    memset(body_f, 0, sizeof body_f);
    if (m->enable_contact) {
        for (int s = 0; s < 2; ++s) {
            int b = 6 + 6 * s;
            double kn = m->contact_k * e->kscale[s], cn = m->contact_c * e->cscale[s];
            for (int c = 0; c < 4; ++c) {
                double rp[3], p[3];
                matvec(k.R[b], m->foot_corner[c], rp);
                for (int r = 0; r < 3; ++r) p[r] = k.x[b][r] + rp[r];
                foot_fn[s] += ground_point(m, &k, b, p, terr, qvel, kn, cn, e->mu[s], M, rhs, body_f);
            }
        }
    }
    /* the other collision primitives against the ground: trunk box corners, lowest rim point of each end
     * cap of the hip-yaw and shank cylinders (resources/T1/T1_locomotion.xml:42,66,71,99,104) */
    if (m->enable_body_contact) {
        for (int c = 0; c < 8; ++c) {
            double l[3], rp[3], p[3];
            for (int r = 0; r < 3; ++r) l[r] = m->trunk_box_pos[r] + (((c >> r) & 1) ? -1.0 : 1.0) * m->trunk_box_half[r];
            matvec(k.R[0], l, rp);
            for (int r = 0; r < 3; ++r) p[r] = k.x[0][r] + rp[r];
            ground_point(m, &k, 0, p, terr, qvel, m->contact_k, m->contact_c, m->body_mu, M, rhs, body_f);
        }
        for (int s = 0; s < 2; ++s)
            for (int ci = 0; ci < 2; ++ci) {
                int b = 1 + 6 * s + 2 + ci;
                double cw[3], a[3] = {k.R[b][2], k.R[b][5], k.R[b][8]}, down[3];
                matvec(k.R[b], m->cyl_pos[ci], cw);
                /* unit vector in the cap plane pointing most steeply down: -(z - (z.a) a) / |.| ; none if the cylinder stands upright */
                double n2 = 1.0 - a[2] * a[2], inv = (n2 > 1e-12) ? 1.0 / sqrt(n2) : 0.0;
                down[0] = a[2] * a[0] * inv; down[1] = a[2] * a[1] * inv; down[2] = -n2 * inv;
                for (int e2 = 0; e2 < 2; ++e2) {
                    double p[3];
                    for (int r = 0; r < 3; ++r)
                        p[r] = k.x[b][r] + cw[r] + (e2 ? -1.0 : 1.0) * m->cyl_half[ci] * a[r] + m->cyl_radius[ci] * down[r];
                    ground_point(m, &k, b, p, terr, qvel, m->contact_k, m->contact_c, m->body_mu, M, rhs, body_f);
                }
            }
    }
    /* leg-leg contacts (asset.self_collisions: 0 = enabled, envs/T1.yaml:69): shank cylinders and foot boxes as capsules.
     * Every (left, right) pair closer than the sum of the radii is pushed apart along the line between the closest axis points:
     * spring on the overlap, damper on the relative normal velocity; linearly implicit per body (block-diagonal: the
     * partner's velocity change is not anticipated - DESIGN.md "contact"). */
    if (m->enable_self_contact) {
        double P[2][2][2][3]; /* [leg][capsule 0 shank / 1 foot][end][xyz], world */
        for (int sd = 0; sd < 2; ++sd) {
            int bs = 1 + 6 * sd + 3, bf = 1 + 6 * sd + 5;
            for (int e2 = 0; e2 < 2; ++e2) {
                double l[3] = {m->cyl_pos[1][0], m->cyl_pos[1][1], m->cyl_pos[1][2] + (e2 ? -1.0 : 1.0) * m->cyl_half[1]}, rp[3];
                matvec(k.R[bs], l, rp);
                for (int r = 0; r < 3; ++r) P[sd][0][e2][r] = k.x[bs][r] + rp[r];
                matvec(k.R[bf], m->foot_cap[e2], rp);
                for (int r = 0; r < 3; ++r) P[sd][1][e2][r] = k.x[bf][r] + rp[r];
            }
        }
        const double ks = m->self_k, cs = m->self_c, dns = cs + dt * ks;
        for (int i = 0; i < 2; ++i)       /* capsule of the left leg */
            for (int j = 0; j < 2; ++j) { /* capsule of the right leg */
                int bl = 1 + (i ? 5 : 3), br = 7 + (j ? 5 : 3);
                double d1[3], d2[3], rr[3];
                for (int r = 0; r < 3; ++r) {
                    d1[r] = P[0][i][1][r] - P[0][i][0][r];
                    d2[r] = P[1][j][1][r] - P[1][j][0][r];
                    rr[r] = P[0][i][0][r] - P[1][j][0][r];
                }
                double a = d1[0] * d1[0] + d1[1] * d1[1] + d1[2] * d1[2], ee = d2[0] * d2[0] + d2[1] * d2[1] + d2[2] * d2[2];
                double f = d2[0] * rr[0] + d2[1] * rr[1] + d2[2] * rr[2], c = d1[0] * rr[0] + d1[1] * rr[1] + d1[2] * rr[2];
                double b = d1[0] * d2[0] + d1[1] * d2[1] + d1[2] * d2[2], den = a * ee - b * b, sp, tp;
                sp = (den > 1e-12 * a * ee) ? fmin(1.0, fmax(0.0, (b * f - c * ee) / den)) : 0.0;
                tp = (b * sp + f) / ee;
                if (tp < 0) { tp = 0; sp = fmin(1.0, fmax(0.0, -c / a)); }
                else if (tp > 1) { tp = 1; sp = fmin(1.0, fmax(0.0, (b - c) / a)); }
                double pl[3], pr[3], dd[3];
                for (int r = 0; r < 3; ++r) {
                    pl[r] = P[0][i][0][r] + sp * d1[r];
                    pr[r] = P[1][j][0][r] + tp * d2[r];
                    dd[r] = pl[r] - pr[r];
                }
                double dist = sqrt(dd[0] * dd[0] + dd[1] * dd[1] + dd[2] * dd[2]);
                double rsum = (i ? m->foot_cap_radius : m->cyl_radius[1]) + (j ? m->foot_cap_radius : m->cyl_radius[1]);
                double depth = rsum - dist;
                if (depth <= 0 || dist <= 1e-6) continue;
                double n[3] = {dd[0] / dist, dd[1] / dist, dd[2] / dist}; /* from the right capsule towards the left one */
                double JvL[3][NV], JwL[3][NV], JvR[3][NV], JwR[3][NV], gl[NV], gr[NV], vn = 0;
                jacobian(&k, bl, pl, JvL, JwL);
                jacobian(&k, br, pr, JvR, JwR);
                for (int q = 0; q < NV; ++q) {
                    gl[q] = n[0] * JvL[0][q] + n[1] * JvL[1][q] + n[2] * JvL[2][q];
                    gr[q] = n[0] * JvR[0][q] + n[1] * JvR[1][q] + n[2] * JvR[2][q];
                    vn += (gl[q] - gr[q]) * qvel[q];
                }
                if (ks * depth - cs * vn <= 0) continue;
                double fmag = ks * depth - dns * vn;
                for (int q = 0; q < NV; ++q) {
                    rhs[q] += (gl[q] - gr[q]) * fmag;
                    for (int q2 = 0; q2 < NV; ++q2) M[q][q2] += dt * dns * (gl[q] * gl[q2] + gr[q] * gr[q2]);
                }
                for (int r = 0; r < 3; ++r) { body_f[3 * bl + r] += fmag * n[r]; body_f[3 * br + r] -= fmag * n[r]; }
            }
    }
    if (m->enable_limits) {
        for (int j = 0; j < 12; ++j) {
            double viol = 0;
            if (e->q[j] < m->jnt_lower[j]) viol = m->jnt_lower[j] - e->q[j];
            else if (e->q[j] > m->jnt_upper[j]) viol = m->jnt_upper[j] - e->q[j];
            if (viol != 0) {
                double ke = m->limit_k * m->dof_inertia[j], ce = m->limit_c * m->dof_inertia[j], de = ce + dt * ke;
                rhs[6 + j] += ke * viol - de * e->qd[j];
                M[6 + j][6 + j] += dt * de;
            }
        }
    }
    if (cholesky_solve(M, rhs) != 0) return -1;
    if (qacc_out) memcpy(qacc_out, rhs, sizeof rhs);
    if (foot_fn_out) { foot_fn_out[0] = foot_fn[0]; foot_fn_out[1] = foot_fn[1]; }
    if (body_f_out) memcpy(body_f_out, body_f, sizeof body_f);
    if (!integrate) return 0;
    /* mj_Euler: velocities first, positions with the NEW velocities, exact quaternion exponential */
    for (int r = 0; r < 3; ++r) {
        e->vlin[r] += dt * rhs[r];
        e->wb[r] += dt * rhs[3 + r];
        e->pos[r] += dt * e->vlin[r];
    }
    for (int j = 0; j < 12; ++j) { e->qd[j] += dt * rhs[6 + j]; e->q[j] += dt * e->qd[j]; }
    {
        double qn = sqrt(e->quat[0] * e->quat[0] + e->quat[1] * e->quat[1] + e->quat[2] * e->quat[2] + e->quat[3] * e->quat[3]);
        double x = e->quat[0] / qn, y = e->quat[1] / qn, z = e->quat[2] / qn, w = e->quat[3] / qn;
        double wn = sqrt(e->wb[0] * e->wb[0] + e->wb[1] * e->wb[1] + e->wb[2] * e->wb[2]);
        double half = 0.5 * dt * wn, kk = (wn > 1e-9) ? sin(half) / wn : 0.5 * dt, dw = cos(half);
        double dx = kk * e->wb[0], dy = kk * e->wb[1], dz = kk * e->wb[2];
        double nq[4] = {w * dx + x * dw + y * dz - z * dy, w * dy - x * dz + y * dw + z * dx,
                        w * dz + x * dy - y * dx + z * dw, w * dw - x * dx - y * dy - z * dz};
        double n2 = sqrt(nq[0] * nq[0] + nq[1] * nq[1] + nq[2] * nq[2] + nq[3] * nq[3]);
        for (int i = 0; i < 4; ++i) e->quat[i] = nq[i] / n2;
    }
    return 0;
}

int t1o_tick(const B200T1ModelD* m, T1OEnv* e, const double* tau, const double* push_f, const double* push_t,
             const T1OTerrain* terr, double* qacc_out, double* foot_fn_out, int integrate) {
    return t1o_tick_f(m, e, tau, push_f, push_t, terr, qacc_out, foot_fn_out, 0, integrate);
}

/* feet world pose (position, quaternion xyzw by the same Shepperd branches as the kernel is NOT required: tests
 * compare rotation matrices) */
void t1o_feet(const B200T1ModelD* m, const T1OEnv* e, double* foot_pos /*2x3*/, double* foot_R /*2x9*/) {
    Kin k;
    kinematics(m, e, &k);
    for (int s = 0; s < 2; ++s) {
        memcpy(foot_pos + 3 * s, k.x[6 + 6 * s], 24);
        memcpy(foot_R + 9 * s, k.R[6 + 6 * s], 72);
    }
}

/* total energy (kinetic + gravitational potential) and linear / angular momentum about the world origin */
void t1o_energy_momentum(const B200T1ModelD* m, const T1OEnv* e, double* energy, double* lin /*3*/, double* ang /*3*/) {
    Kin k;
    kinematics(m, e, &k);
    double qvel[NV], E = 0, P[3] = {0, 0, 0}, L[3] = {0, 0, 0};
    for (int r = 0; r < 3; ++r) { qvel[r] = e->vlin[r]; qvel[3 + r] = e->wb[r]; }
    for (int j = 0; j < 12; ++j) qvel[6 + j] = e->qd[j];
    for (int b = 0; b < NB; ++b) {
        double Jv[3][NV], Jw[3][NV], v[3] = {0, 0, 0}, w[3] = {0, 0, 0}, Iw_w[3], t[3];
        jacobian(&k, b, k.c[b], Jv, Jw);
        for (int r = 0; r < 3; ++r)
            for (int i = 0; i < NV; ++i) { v[r] += Jv[r][i] * qvel[i]; w[r] += Jw[r][i] * qvel[i]; }
        matvec(k.Iw[b], w, Iw_w);
        E += 0.5 * e->mass[b] * (v[0] * v[0] + v[1] * v[1] + v[2] * v[2]) + 0.5 * (w[0] * Iw_w[0] + w[1] * Iw_w[1] + w[2] * Iw_w[2]);
        E += e->mass[b] * m->gravity * k.c[b][2];
        for (int r = 0; r < 3; ++r) P[r] += e->mass[b] * v[r];
        double mv[3] = {e->mass[b] * v[0], e->mass[b] * v[1], e->mass[b] * v[2]};
        cross(k.c[b], mv, t);
        for (int r = 0; r < 3; ++r) L[r] += t[r] + Iw_w[r];
    }
    *energy = E;
    memcpy(lin, P, 24);
    memcpy(ang, L, 24);
}

/* ---- the decimated PD loop of one env.step (envs/t1.py:439-456) for `nenv` independent envs ---------------------
 * actions [nenv][12] (already clipped), kp/kd/fric [nenv][12], delay [nenv]; last_targets in/out [nenv][12];
 * torques_mean out [nenv][12]. mjcf_mode & 1: play_mujoco.py:751-755 (no delay, no joint friction).  The push acts on the
 * first substep only (Isaac Gym applies the force tensors of envs/t1.py:522-527 to the next simulate() call); mjcf_mode & 2: on
 * every substep. */
int t1o_env_physics(const B200T1ModelD* m, T1OEnv* envs, int nenv, const double* actions, const double* default_q,
                    double action_scale, const double* kp, const double* kd, const double* fric,
                    const double* torque_limit, const int* delay, double* last_targets, const double* push_f,
                    const double* push_t, const T1OTerrain* terr, int decimation, double* torques_mean, int mjcf_mode) {
    int bad = 0;
    const int push_all = (mjcf_mode & 2) != 0;
    const double zero3[3] = {0.0, 0.0, 0.0};
    mjcf_mode &= 1;
#pragma omp parallel for schedule(static) reduction(+ : bad)
    for (int n = 0; n < nenv; ++n) {
        double tgt[12], tau[12], acc[12];
        for (int j = 0; j < 12; ++j) { tgt[j] = default_q[j] + action_scale * actions[12 * n + j]; acc[j] = 0; }
        for (int i = 0; i < decimation; ++i) {
            if (mjcf_mode || delay[n] == i)
                for (int j = 0; j < 12; ++j) last_targets[12 * n + j] = tgt[j];
            for (int j = 0; j < 12; ++j) {
                double t = kp[12 * n + j] * (last_targets[12 * n + j] - envs[n].q[j]) - kd[12 * n + j] * envs[n].qd[j];
                if (!mjcf_mode) {
                    double f = fmin(fric[12 * n + j], fabs(t));
                    t -= (t > 0 ? f : (t < 0 ? -f : 0.0));
                }
                t = fmax(-torque_limit[j], fmin(torque_limit[j], t));
                tau[j] = t;
                acc[j] += t;
            }
            const int pushed = push_all || i == 0;
            if (t1o_tick(m, &envs[n], tau, pushed ? push_f + 3 * n : zero3, pushed ? push_t + 3 * n : zero3, terr, 0, 0, 1) != 0) bad += 1;
        }
        for (int j = 0; j < 12; ++j) torques_mean[12 * n + j] = acc[j] / decimation;
    }
    return bad;
}

int t1o_sizeof_env(void) { return (int)sizeof(T1OEnv); }
