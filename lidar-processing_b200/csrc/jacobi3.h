// 3x3 one-thread Jacobi SVD used by the plane fit, host- and device-compilable.
//
// The reference's plane normal is `svd_solver_.matrixV().col(2)` of an
// Eigen::JacobiSVD<Eigen::Matrix3f>(ComputeThinV) (reference src/segmentation.hpp:110,
// src/segmentation.cpp:87-94). The ground test is a SIGNED distance (segmentation.cpp:290-299), so
// the sign Eigen's two-sided Jacobi iteration leaves on that column matters; a closed-form
// eigensolver would lose it. This routine follows Eigen 3.4's published procedure for the square
// real case: scale by max|a_ij|; sweeps over (p,q) = (1,0),(2,0),(2,1); each 2x2 block is first
// symmetrised by a rotation, then diagonalised by a Jacobi rotation; V accumulates the right
// rotations and is never sign-flipped; columns are finally ordered by descending singular value.
//
// Must be compiled without FMA contraction (nvcc -fmad=false; x86-64 baseline g++ has no FMA), like
// the reference build.
#pragma once

#include <math.h>

#ifdef __CUDACC__
#define LB_J3_HD __host__ __device__ inline
#else
#define LB_J3_HD inline
#endif

namespace lb
{

struct J3Rot
{
    float c;
    float s;
};

LB_J3_HD void j3_rotate(float &x, float &y, float c, float s)
{
    const float xi = x;
    const float yi = y;
    x = c * xi + s * yi;
    y = -s * xi + c * yi;
}

// FLT_MIN / FLT_EPSILON spelled out so the header needs no <cfloat> on device
#define LB_J3_FLT_MIN 1.17549435e-38f
#define LB_J3_FLT_EPS 1.19209290e-07f

// a: row-major symmetric 3x3. v: row-major 3x3, columns = right singular vectors sorted by
// descending singular value; sv: singular values. Returns false when the input is not finite
// (Eigen: info() == InvalidInput -> the reference treats the fit as failed, segmentation.cpp:88-92).
LB_J3_HD bool jacobi_svd3(const float a[9], float v[9], float sv[3])
{
    float scale = 0.0f;
    for (int i = 0; i < 9; ++i)
    {
        const float m = fabsf(a[i]);
        if (!(m <= scale))
            scale = m;
    }
    if (!(fabsf(scale) <= 3.402823466e+38f)) // NaN or Inf
        return false;
    if (scale == 0.0f)
        scale = 1.0f;

    float w[9];
    for (int i = 0; i < 9; ++i)
    {
        w[i] = a[i] / scale;
        v[i] = (i == 0 || i == 4 || i == 8) ? 1.0f : 0.0f;
    }
    float max_diag = fmaxf(fabsf(w[0]), fmaxf(fabsf(w[4]), fabsf(w[8])));
    const float precision = 2.0f * LB_J3_FLT_EPS;

    bool finished = false;
    int sweeps = 0;
    while (!finished && sweeps < 64)
    {
        finished = true;
        ++sweeps;
        for (int p = 1; p < 3; ++p)
        {
            for (int q = 0; q < p; ++q)
            {
                const float threshold = fmaxf(LB_J3_FLT_MIN, precision * max_diag);
                if (!(fabsf(w[p * 3 + q]) > threshold || fabsf(w[q * 3 + p]) > threshold))
                    continue;
                finished = false;

                // 2x2 block [[m00 m01],[m10 m11]] = rows/cols (p,q)
                float m00 = w[p * 3 + p], m01 = w[p * 3 + q], m10 = w[q * 3 + p], m11 = w[q * 3 + q];
                J3Rot rot1;
                const float t = m00 + m11;
                const float d = m10 - m01;
                if (fabsf(d) < LB_J3_FLT_MIN)
                {
                    rot1.s = 0.0f;
                    rot1.c = 1.0f;
                }
                else
                {
                    const float u = t / d;
                    const float tmp = sqrtf(1.0f + u * u);
                    rot1.s = 1.0f / tmp;
                    rot1.c = u / tmp;
                }
                if (!(rot1.c == 1.0f && rot1.s == 0.0f))
                {
                    j3_rotate(m00, m10, rot1.c, rot1.s);
                    j3_rotate(m01, m11, rot1.c, rot1.s);
                }
                // right rotation diagonalising the (now symmetric) block
                J3Rot jr;
                const float deno = 2.0f * fabsf(m01);
                if (deno < LB_J3_FLT_MIN)
                {
                    jr.c = 1.0f;
                    jr.s = 0.0f;
                }
                else
                {
                    const float tau = (m00 - m11) / deno;
                    const float ww = sqrtf(tau * tau + 1.0f);
                    const float tt = tau > 0.0f ? 1.0f / (tau + ww) : 1.0f / (tau - ww);
                    const float sign_t = tt > 0.0f ? 1.0f : -1.0f;
                    const float n = 1.0f / sqrtf(tt * tt + 1.0f);
                    jr.s = -sign_t * (m01 / fabsf(m01)) * fabsf(tt) * n;
                    jr.c = n;
                }
                // left rotation = rot1 * transpose(jr)
                J3Rot jl;
                jl.c = rot1.c * jr.c - rot1.s * (-jr.s);
                jl.s = rot1.c * (-jr.s) + rot1.s * jr.c;

                if (!(jl.c == 1.0f && jl.s == 0.0f))
                    for (int i = 0; i < 3; ++i)
                        j3_rotate(w[p * 3 + i], w[q * 3 + i], jl.c, jl.s); // rows p,q
                // applyOnTheRight(p,q,jr) rotates columns p,q with transpose(jr) = (c, -s)
                if (!(jr.c == 1.0f && -jr.s == 0.0f))
                    for (int i = 0; i < 3; ++i)
                    {
                        j3_rotate(w[i * 3 + p], w[i * 3 + q], jr.c, -jr.s);
                        j3_rotate(v[i * 3 + p], v[i * 3 + q], jr.c, -jr.s);
                    }
                max_diag = fmaxf(max_diag, fmaxf(fabsf(w[p * 3 + p]), fabsf(w[q * 3 + q])));
            }
        }
    }

    for (int i = 0; i < 3; ++i)
        sv[i] = fabsf(w[i * 3 + i]) * scale;
    for (int i = 0; i < 3; ++i)
    {
        int pos = i;
        float best = sv[i];
        for (int k = i + 1; k < 3; ++k)
            if (sv[k] > best)
            {
                best = sv[k];
                pos = k;
            }
        if (best == 0.0f)
            break;
        if (pos != i)
        {
            const float ts = sv[i];
            sv[i] = sv[pos];
            sv[pos] = ts;
            for (int r = 0; r < 3; ++r)
            {
                const float tv = v[r * 3 + i];
                v[r * 3 + i] = v[r * 3 + pos];
                v[r * 3 + pos] = tv;
            }
        }
    }
    return true;
}

} // namespace lb
