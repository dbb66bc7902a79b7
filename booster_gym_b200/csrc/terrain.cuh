// terrain.cuh - Terrain.terrain_heights (utils/terrain.py:101-121) on a device-resident int16 heightfield.
//
// Bit-exact restatement of the reference's NumPy arithmetic: the cell coordinate is formed in fp32
// (`border_pixels + fp32(px) / horizontal_scale`, NumPy 2 keeps the Python float weak -> fp32 division, which must be
// a correctly rounded IEEE division: multiplying by 10 moves ~6.6 ppm of positions into the neighbouring cell,
// SURVEY 7 hard part 6), floor -> int64, then the bilinear weights and the 4-tap sum run in fp64 in the reference's
// left-to-right order with no fused multiply-add, times vertical_scale (Python float = fp64), cast to fp32.
// NumPy wraps negative indices (no clamp in the reference); indices beyond the array raise IndexError there and
// return 0 here (robots are teleported back long before, envs/t1.py:343-360).
#pragma once
#include <math.h>
#include <stdint.h>

#include "rng.cuh"  // B200_HD / B200_HD_CALL

#if defined(__CUDACC__)
#define B200_HD __host__ __device__ __forceinline__
#else
#ifndef B200_HD
#define B200_HD inline
#endif
#endif

namespace b200 {

struct TerrainView {
    const int16_t* hf;  // nullptr: plane (utils/terrain.py:102-103 -> zeros)
    int rows, cols, border_pixels;
    float horizontal_scale;
    double vertical_scale;
    float max_height = 3.0e38f;  // upper bound of every lookup (culls the rarely touching collision shapes); default: never cull

    B200_HD float operator()(float px, float py) const {
        if (hf == nullptr) return 0.0f;
        return lookup(px, py);
    }
    B200_HD_CALL float lookup(float px, float py) const {
#if defined(__CUDA_ARCH__)
        const float x = __fadd_rn((float)border_pixels, __fdiv_rn(px, horizontal_scale));
        const float y = __fadd_rn((float)border_pixels, __fdiv_rn(py, horizontal_scale));
#else
        const volatile float qx = px / horizontal_scale, qy = py / horizontal_scale;
        const float x = (float)border_pixels + qx, y = (float)border_pixels + qy;
#endif
        const long long x1 = (long long)floorf(x), y1 = (long long)floorf(y);
        const long long x2 = x1 + 1, y2 = y1 + 1;
        long long ix1 = x1 < 0 ? x1 + rows : x1, ix2 = x2 < 0 ? x2 + rows : x2;
        long long iy1 = y1 < 0 ? y1 + cols : y1, iy2 = y2 < 0 ? y2 + cols : y2;
        if (ix1 < 0 || ix1 >= rows || ix2 < 0 || ix2 >= rows || iy1 < 0 || iy1 >= cols || iy2 < 0 || iy2 >= cols) return 0.0f;
        const double dx2 = (double)x2 - (double)x, dx1 = (double)x - (double)x1;
        const double dy2 = (double)y2 - (double)y, dy1 = (double)y - (double)y1;
        const double h11 = (double)hf[ix1 * cols + iy1], h21 = (double)hf[ix2 * cols + iy1];
        const double h12 = (double)hf[ix1 * cols + iy2], h22 = (double)hf[ix2 * cols + iy2];
#if defined(__CUDA_ARCH__)
        double acc = __dmul_rn(__dmul_rn(dx2, dy2), h11);
        acc = __dadd_rn(acc, __dmul_rn(__dmul_rn(dx1, dy2), h21));
        acc = __dadd_rn(acc, __dmul_rn(__dmul_rn(dx2, dy1), h12));
        acc = __dadd_rn(acc, __dmul_rn(__dmul_rn(dx1, dy1), h22));
        return (float)__dmul_rn(acc, vertical_scale);
#else
        volatile double t1 = dx2 * dy2, t2 = dx1 * dy2, t3 = dx2 * dy1, t4 = dx1 * dy1;
        volatile double p1 = t1 * h11, p2 = t2 * h21, p3 = t3 * h12, p4 = t4 * h22;
        volatile double acc = p1 + p2;
        acc = acc + p3;
        acc = acc + p4;
        volatile double r = acc * vertical_scale;
        return (float)r;
#endif
    }
};

}  // namespace b200
