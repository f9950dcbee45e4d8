// common.cuh -- shared device helpers of the B200 RANSAC voting layer.
//
// Exact arithmetic: every floating-point operation the reference kernels perform is
// written with round-to-nearest intrinsics in the contraction pattern nvcc 12.9 gives
// the reference source for sm_100 (read from `cuobjdump -sass` of the unmodified
// reference build, see DESIGN.md "Observed arithmetic").  That is what makes hypotheses
// bit-equal and inlier counts equal to the reference extension.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace pvb {

// ---------------------------------------------------------------------------------
// Philox4x32-10.  Counter layout of the built-in sampling mode (DESIGN.md "Sampling"):
//   pair indices : ctr = (h, k, image, tag_idx)  -> t0 = out[0] % tn, t1 = out[1] % tn
//   thinning     : ctr = (pixel>>2, 0, image, tag_sel) -> u = (out[pixel&3] >> 8) * 2^-24
// Keyed by (seed_lo, seed_hi).  Results do not depend on grid shape or GPU count.
// ---------------------------------------------------------------------------------
__device__ __forceinline__ uint4 philox4x32_10(uint4 c, uint2 k)
{
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        const uint32_t hi0 = __umulhi(0xD2511F53u, c.x), lo0 = 0xD2511F53u * c.x;
        const uint32_t hi1 = __umulhi(0xCD9E8D57u, c.z), lo1 = 0xCD9E8D57u * c.z;
        c = make_uint4(hi1 ^ c.y ^ k.x, lo1, hi0 ^ c.w ^ k.y, lo0);
        k.x += 0x9E3779B9u;
        k.y += 0xBB67AE85u;
    }
    return c;
}

__device__ __forceinline__ float u32_to_unit(uint32_t x) { return (float)(x >> 8) * 5.9604644775390625e-08f; }

// ---------------------------------------------------------------------------------
// Hypothesis from one pixel pair (reference: ransac_voting_kernel.cu:27-48).
// Returns false where the reference thread returns early; the caller writes (0,0).
// ---------------------------------------------------------------------------------
__device__ __forceinline__ bool hypothesis_from_pair(float dx0, float dy0, float cx0, float cy0,
                                                     float dx1, float dy1, float cx1, float cy1,
                                                     float &x, float &y)
{
    const float p = __fmul_rn(dy0, dx1);
    const float q = __fmul_rn(dx0, dy1);
    const float det1 = __fsub_rn(p, q);   // nx1*ny0 - nx0*ny1   (.cu:42)
    const float det2 = __fsub_rn(q, p);   // ny1*nx0 - ny0*nx1   (.cu:43)
    if (fabs((double)det1) < 1e-6) return false;
    if (fabs((double)det2) < 1e-6) return false;
    const float e0 = __fmaf_rn(dy0, cx0, -__fmul_rn(dx0, cy0));
    const float e1 = __fmaf_rn(dy1, cx1, -__fmul_rn(dx1, cy1));
    y = __fdiv_rn(__fmaf_rn(dy1, e0, -__fmul_rn(dy0, e1)), det1);   // .cu:44
    x = __fdiv_rn(__fmaf_rn(dx0, e1, -__fmul_rn(dx1, e0)), det2);   // .cu:45
    return true;
}

// ---------------------------------------------------------------------------------
// The reference inlier predicate, exactly (ransac_voting_kernel.cu:107-125).
// ---------------------------------------------------------------------------------
__device__ __forceinline__ bool vote_exact(float vx, float vy, float cx, float cy, float hx, float hy,
                                           float thresh)
{
    const float dx = __fsub_rn(hx, cx), dy = __fsub_rn(hy, cy);
    const float n1sq = __fmaf_rn(vx, vx, __fmul_rn(vy, vy));
    const float n2sq = __fmaf_rn(dx, dx, __fmul_rn(dy, dy));
    const float norm1 = __fsqrt_rn(n1sq), norm2 = __fsqrt_rn(n2sq);
    if ((double)norm1 < 1e-6 || (double)norm2 < 1e-6) return false;
    const float den = __fmul_rn(norm2, norm1);
    const float dot = __fmaf_rn(vx, dx, __fmul_rn(vy, dy));
    return __fdiv_rn(dot, den) > thresh;
}

// Largest float below 1e-6: (double)n < 1e-6  <=>  n <= below_1e6() for a float n
// (float(1e-6) = 0x358637BD = 9.99999997e-07 < 1e-6 < next float).
__device__ __forceinline__ float below_1e6() { return __int_as_float(0x358637BD); }

__device__ __forceinline__ float warp_sum(float v)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ double warp_sum(double v)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ int warp_sum(int v) { return __reduce_add_sync(0xffffffffu, v); }

// inv(sqrtm(cov)) of one 2x2 covariance, packed (wxx, wxy, wyy) and rounded to fp32 -- SURVEY 8f row 2
// (lib/evaluators/linemod/pvnet.py:118-130: scipy.linalg.sqrtm + np.linalg.inv per keypoint on the CPU).
// Closed form for a symmetric positive definite 2x2 M: sqrt(M) = (M + s I)/t, s = sqrt(det M), t = sqrt(tr M + 2 s),
// so inv(sqrt(M)) = t * adj(M + s I) / det(M + s I).  cov[0,0] < 1e-6 or any NaN -> zeros, like the reference.
__device__ __forceinline__ void cov_to_weights(const float4 c, float &o0, float &o1, float &o2)
{
    o0 = 0.f; o1 = 0.f; o2 = 0.f;
    const bool bad = (c.x < 1e-6f) || (c.x != c.x) || (c.y != c.y) || (c.z != c.z) || (c.w != c.w);
    if (!bad) {
        const double a = c.x, b = 0.5 * ((double)c.y + (double)c.z), d = c.w;
        const double det = a * d - b * b;
        if (det > 0.0) {
            const double s = sqrt(det), t = sqrt(a + d + 2.0 * s);
            const double a2 = a + s, d2 = d + s;
            const double den = a2 * d2 - b * b;
            o0 = (float)(t * d2 / den); o1 = (float)(-t * b / den); o2 = (float)(t * a2 / den);
        }
    }
}

// Parameters of the cone test used by the fast path of the vote kernel (vote.cu).
struct ConeParams {
    float kappa;   // tan(acos(thresh)) = sqrt(1-t^2)/t
    float band;    // guard band per unit of S = |hx-ox|+|hy-oy|+cmax(tile) ; +inf => exact path only
    float thresh;  // (float)inlier_thresh
};

} // namespace pvb
