// pnp.cu -- batched uncertainty-PnP refinement: one warp per pose problem.
//
// Replaces the per-image CPU call lib/csrc/uncertainty_pnp/src/uncertainty_pnp.cpp:61-92 (ceres::Solve on 6 parameters and
// 2*pn residuals, pn = 9..17 keypoints).  A problem is far too small for more than a warp: lane i owns points i, i+32, ...,
// evaluates their residuals and 2x6 Jacobians (pnp_core.cuh) and the warp adds the 28 numbers of the normal equations with
// an XOR butterfly (every lane ends with the same bits, so all lanes run the identical trust-region state machine without
// divergence or broadcasts).  Everything is fp64 like the reference.  Latency-bound by design: ~10 evaluations per problem.
#include "common.cuh"
#include "kernels.h"
#include "pnp_core.cuh"

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

cudaError_t launch_pnp(const PnpArgs &a, cudaStream_t st)
{
    if (a.n <= 0) return cudaSuccess;
    pnp_kernel<<<(a.n + 3) / 4, 128, 0, st>>>(a);
    return cudaGetLastError();
}

} // namespace pvb
