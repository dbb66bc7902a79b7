// rng.cuh - counter-based random numbers for the in-kernel domain randomisation / noise.
//
// The reference draws from torch's global generator (torch.randn_like / rand_like / randint, utils/utils.py:11,15,
// envs/t1.py:316,335,384) - a sequential stream that cannot be reproduced per environment on the device.  Here every
// draw is Philox4x32-10 of the counter (global env index, step, purpose, sub) under the key (seed), so a sample depends
// only on WHAT it is for, never on launch geometry or on which other envs reset.  Parity tests read the exact samples
// back through b200_rng_fill() and inject them into the reference (monkey-patched torch.randn_like), so the
// comparison is on identical noise.
#pragma once
#include <math.h>
#include <stdint.h>

#if defined(__CUDACC__)
#define B200_HD __host__ __device__ __forceinline__
#else
#ifndef B200_HD
#define B200_HD inline
#endif
#endif
// Large helpers that are called many times per thread are kept OUT of line on the device: the env kernels execute each
// instruction once per warp with 1-4 warps per SM, so they are bound by serialised instruction-cache misses
// (ncu: stall_no_instruction dominant); a called function is fetched once and then hits the i-cache.
#if defined(__CUDACC__)
#define B200_HD_CALL inline __host__ __device__ __noinline__
#else
#define B200_HD_CALL inline
#endif

namespace b200 {

enum RngPurpose {
    RP_INIT_GAINS = 1,   // sub 0-2 kp, 3-5 kd, 6-8 friction            (uniform/gaussian per config)
    RP_INIT_BODY = 2,    // sub 0: base com xyz + mass; sub 1+3b.. other bodies (sub = 1 + b: com xyz + mass)
    RP_INIT_FOOT = 3,    // sub s: friction, compliance, restitution of foot s
    RP_RESET_DOF = 4,    // sub 0-2: 12 values
    RP_RESET_ROOT = 5,   // sub 0: x, y, yaw(uniform always), -; sub 1: vx, vy
    RP_RESET_DELAY = 6,  // sub 0 word 0: delay_steps
    RP_COMMAND = 7,      // sub 0: vx, vy, yaw, gait (uniform); sub 1: word0 still draw, word1 resample interval
    RP_KICK = 8,         // sub 0: lin xyz; sub 1: ang xyz
    RP_PUSH = 9,         // sub 0: force xyz; sub 1: torque xyz
    RP_OBS_NOISE = 10,   // 34 values, sub = i / 4: gravity 0-2, ang_vel 3-5, dof_pos 6-17, dof_vel 18-29, lin_vel 30-32, height 33
    RP_POLICY = 11       // action sampling noise: sub 0-2 (12 normals)
};

#define B200_RNG_SHARED_ENV 0xFFFFFFFFu  /* env id of draws shared by all envs of a call (reset dof noise) */

struct Philox4 {
    uint32_t w[4];
};

B200_HD uint32_t mulhi32(uint32_t a, uint32_t b) {
#if defined(__CUDA_ARCH__)
    return __umulhi(a, b);
#else
    return (uint32_t)(((uint64_t)a * (uint64_t)b) >> 32);
#endif
}

B200_HD_CALL Philox4 philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0, uint32_t k1) {
    const uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u, W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        const uint32_t hi0 = mulhi32(M0, c0), lo0 = M0 * c0;
        const uint32_t hi1 = mulhi32(M1, c2), lo1 = M1 * c2;
        const uint32_t n0 = hi1 ^ c1 ^ k0, n2 = hi0 ^ c3 ^ k1;
        c0 = n0; c1 = lo1; c2 = n2; c3 = lo0;
        k0 += W0; k1 += W1;
    }
    Philox4 o;
    o.w[0] = c0; o.w[1] = c1; o.w[2] = c2; o.w[3] = c3;
    return o;
}

// the one place that defines the counter layout
B200_HD Philox4 rng_words(uint64_t seed, uint32_t env_global, uint64_t step, int purpose, int sub) {
    return philox4x32_10(env_global, (uint32_t)step, (uint32_t)(step >> 32) ^ ((uint32_t)purpose << 16), (uint32_t)sub,
                         (uint32_t)seed, (uint32_t)(seed >> 32));
}

B200_HD float u01(uint32_t w) { return (float)(w >> 8) * (1.0f / 16777216.0f); }            // [0, 1)
B200_HD float u01_open(uint32_t w) { return ((float)(w >> 8) + 1.0f) * (1.0f / 16777216.0f); }  // (0, 1]

struct Rand4 {
    float u[4];  // uniform [0,1)
    float n[4];  // standard normal (Box-Muller on word pairs)
};

B200_HD_CALL Rand4 rand4(const Philox4& p) {
    Rand4 r;
#pragma unroll
    for (int i = 0; i < 4; ++i) r.u[i] = u01(p.w[i]);
#pragma unroll
    for (int i = 0; i < 2; ++i) {
        const float rad = sqrtf(-2.0f * logf(u01_open(p.w[2 * i])));
        const float ang = 6.28318530717958647692f * u01(p.w[2 * i + 1]);
        r.n[2 * i] = rad * cosf(ang);
        r.n[2 * i + 1] = rad * sinf(ang);
    }
    return r;
}

}  // namespace b200
