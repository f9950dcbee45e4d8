// pnp.cu -- batched uncertainty-PnP refinement: one warp per pose problem.
//
// Replaces the per-image CPU call lib/csrc/uncertainty_pnp/src/uncertainty_pnp.cpp:61-92 (ceres::Solve on 6 parameters and
// 2*pn residuals, pn = 9..17 keypoints).  A problem is far too small for more than a warp: lane i owns points i, i+32, ...,
// evaluates their residuals and 2x6 Jacobians (pnp_core.cuh) and the warp adds the 28 numbers of the normal equations with
// an XOR butterfly (every lane ends with the same bits, so all lanes run the identical trust-region state machine without
// divergence or broadcasts).  Everything is fp64 like the reference.  Latency-bound by design: ~10 evaluations per problem.
#include <math_constants.h>
#include "common.cuh"
#include "kernels.h"
#include "pnp_core.cuh"
#include "p3p_core.cuh"

namespace pvb {

__device__ __forceinline__ void pnp_warp_normal(const double *pose, const PnpArgs &a, int prob, int lane, PnpNormal &n)
{
    const double *p2 = a.pts2d + (size_t)prob * a.pn * 2;
    const double *p3 = a.pts3d + (size_t)prob * a.pts3d_stride;
    const double *w = a.wgt2d + (size_t)prob * a.pn * 3;
    const double *K = a.K + (size_t)prob * a.k_stride;
    const double cam[4] = { K[0], K[4], K[2], K[5] };     // fx, fy, px, py (uncertainty_pnp.cpp:77)
    pnp_normal_zero(n);
    for (int i = lane; i < a.pn; i += 32) pnp_accumulate_point(pose, p3 + 3 * i, p2 + 2 * i, w + 3 * i, cam, n);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
#pragma unroll
        for (int q = 0; q < 21; ++q) n.H[q] += __shfl_xor_sync(0xffffffffu, n.H[q], o);
#pragma unroll
        for (int q = 0; q < 6; ++q) n.g[q] += __shfl_xor_sync(0xffffffffu, n.g[q], o);
        n.cost += __shfl_xor_sync(0xffffffffu, n.cost, o);
    }
}

__global__ void __launch_bounds__(128)
pnp_kernel(PnpArgs a)
{
    const int lane = threadIdx.x & 31;
    const int prob = blockIdx.x * 4 + (threadIdx.x >> 5);
    if (prob >= a.n) return;
    PnpOptions opt;
    opt.max_num_iterations = a.max_num_iterations; opt.function_tolerance = a.function_tolerance;
    opt.gradient_tolerance = a.gradient_tolerance; opt.parameter_tolerance = a.parameter_tolerance;
    double init[6];
#pragma unroll
    for (int i = 0; i < 6; ++i) init[i] = a.init_rt[(size_t)prob * 6 + i];
    PnpState st;
    PnpNormal n;
    pnp_warp_normal(init, a, prob, lane, n);
    pnp_init(st, init, n);
    double cand[6];
    while (pnp_propose(st, opt, cand)) {
        pnp_warp_normal(cand, a, prob, lane, n);
        if (!pnp_update(st, opt, cand, n)) break;
    }
    if (lane < 6) a.result_rt[(size_t)prob * 6 + lane] = st.x[lane];
    if (a.info && lane == 0) { a.info[2 * prob] = st.iterations; a.info[2 * prob + 1] = st.code; }
}

// Initial poses: the reference's recipe (un_pnp_utils.py:25-31) -- the four keypoints with the largest wxx + wxy, ascending;
// P3P on the first three, the best-weighted one picks the root (p3p_core.cuh).  One thread per problem: the work is a
// short scalar chain (quartic, <= 4 alignments).  No admissible solution -> NaN pose, which is what OpenCV hands the
// reference in that case; the refinement then stops at once with that pose.
__global__ void __launch_bounds__(64)
p3p_init_kernel(PnpArgs a)
{
    const int prob = blockIdx.x * 64 + threadIdx.x;
    if (prob >= a.n) return;
    const double *p2 = a.pts2d + (size_t)prob * a.pn * 2;
    const double *p3 = a.pts3d + (size_t)prob * a.pts3d_stride;
    const double *w = a.wgt2d + (size_t)prob * a.pn * 3;
    const double *K = a.K + (size_t)prob * a.k_stride;
    const double cam[4] = { K[0], K[4], K[2], K[5] };
    int idx[4];
    p3p_select4(w, a.pn, idx);
    double X[4][3], x2[4][2], rt[6];
    const double nan = CUDART_NAN;
#pragma unroll
    for (int i = 0; i < 6; ++i) rt[i] = nan;
    bool ok = a.pn >= 4;
    if (ok) {
        for (int r = 0; r < 4; ++r) {
            for (int c = 0; c < 3; ++c) X[r][c] = p3[3 * idx[r] + c];
            x2[r][0] = p2[2 * idx[r]]; x2[r][1] = p2[2 * idx[r] + 1];
        }
        p3p_solve4(X, x2, cam, rt);
    }
#pragma unroll
    for (int i = 0; i < 6; ++i) a.result_rt[(size_t)prob * 6 + i] = rt[i];
}

cudaError_t launch_p3p_init(const PnpArgs &a, cudaStream_t st)
{
    if (a.n <= 0) return cudaSuccess;
    p3p_init_kernel<<<(a.n + 63) / 64, 64, 0, st>>>(a);
    return cudaGetLastError();
}

cudaError_t launch_pnp(const PnpArgs &a, cudaStream_t st)
{
    if (a.n <= 0) return cudaSuccess;
    pnp_kernel<<<(a.n + 3) / 4, 128, 0, st>>>(a);
    return cudaGetLastError();
}

} // namespace pvb
