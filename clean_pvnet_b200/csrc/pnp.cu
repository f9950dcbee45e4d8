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

// normal equations of one problem at `pose`: lane i owns points i, i+32, ...; butterfly sum -> every lane holds the same bits
__device__ __forceinline__ void pnp_warp_normal_at(const double *pose, const double *p2, const double *p3, const double *w,
                                                   const double *cam, int pn, int lane, PnpNormal &n)
{
    pnp_normal_zero(n);
    for (int i = lane; i < pn; i += 32) pnp_accumulate_point(pose, p3 + 3 * i, p2 + 2 * i, w + 3 * i, cam, n);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
#pragma unroll
        for (int q = 0; q < 21; ++q) n.H[q] += __shfl_xor_sync(0xffffffffu, n.H[q], o);
#pragma unroll
        for (int q = 0; q < 6; ++q) n.g[q] += __shfl_xor_sync(0xffffffffu, n.g[q], o);
        n.cost += __shfl_xor_sync(0xffffffffu, n.cost, o);
    }
}

__device__ __forceinline__ void pnp_warp_normal(const double *pose, const PnpArgs &a, int prob, int lane, PnpNormal &n)
{
    const double *K = a.K + (size_t)prob * a.k_stride;
    const double cam[4] = { K[0], K[4], K[2], K[5] };     // fx, fy, px, py (uncertainty_pnp.cpp:77)
    pnp_warp_normal_at(pose, a.pts2d + (size_t)prob * a.pn * 2, a.pts3d + (size_t)prob * a.pts3d_stride,
                       a.wgt2d + (size_t)prob * a.pn * 3, cam, a.pn, lane, n);
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

// ---------------------------------------------------------------------------------------------------------------------
// The un_pnp tail in ONE launch (SURVEY 8f rows 2+3 fused): what lib/evaluators/linemod/pvnet.py:118-130 +
// un_pnp_utils.uncertainty_pnp (:6-57) do per image on the CPU -- weights = inv(sqrtm(cov)) per keypoint, P3P initial pose
// on the four best-weighted keypoints, Ceres refinement -- straight from the fp32 tensors the voting layer produced
// (kpt_2d [n,pn,2], var [n,pn,2,2]).  One warp per problem; the problem's data is converted once into shared memory
// (fp64, weights rounded to fp32 first so the result is bit-identical to pvb_uncertainty_weights -> pvb_uncertainty_pnp_init
// -> pvb_uncertainty_pnp run one after the other); lane 0 runs the scalar P3P chain, then all lanes run the LM loop.
// ---------------------------------------------------------------------------------------------------------------------
constexpr int PNPF_WARPS = 4;

__global__ void __launch_bounds__(PNPF_WARPS * 32)
pnp_fused_kernel(PnpFusedArgs a)
{
    __shared__ double s_p2[PNPF_WARPS][PNP_FUSED_MAX_PN * 2], s_p3[PNPF_WARPS][PNP_FUSED_MAX_PN * 3],
        s_w[PNPF_WARPS][PNP_FUSED_MAX_PN * 3];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const int prob = blockIdx.x * PNPF_WARPS + wid;
    if (prob >= a.n) return;
    double *p2 = s_p2[wid], *p3 = s_p3[wid], *w = s_w[wid];
    const double *K = a.K + (size_t)prob * a.k_stride;
    const double cam[4] = { K[0], K[4], K[2], K[5] };
    const double *X = a.pts3d + (size_t)prob * a.pts3d_stride;
    for (int i = lane; i < a.pn; i += 32) {
        const float2 q = __ldg(reinterpret_cast<const float2 *>(a.kpt2d) + (size_t)prob * a.pn + i);
        p2[2 * i] = (double)q.x; p2[2 * i + 1] = (double)q.y;
        float w0, w1, w2;
        if (a.cov) cov_to_weights(__ldg(reinterpret_cast<const float4 *>(a.cov) + (size_t)prob * a.pn + i), w0, w1, w2);
        else { const float *ww = a.weights + ((size_t)prob * a.pn + i) * 3; w0 = ww[0]; w1 = ww[1]; w2 = ww[2]; }
        w[3 * i] = (double)w0; w[3 * i + 1] = (double)w1; w[3 * i + 2] = (double)w2;
        if (a.weights_out) { float *wo = a.weights_out + ((size_t)prob * a.pn + i) * 3; wo[0] = w0; wo[1] = w1; wo[2] = w2; }
        p3[3 * i] = X[3 * i]; p3[3 * i + 1] = X[3 * i + 1]; p3[3 * i + 2] = X[3 * i + 2];
    }
    __syncwarp();
    double init[6];
    if (a.init_rt) {
#pragma unroll
        for (int i = 0; i < 6; ++i) init[i] = a.init_rt[(size_t)prob * 6 + i];
    } else {
        const double nan = CUDART_NAN;
#pragma unroll
        for (int i = 0; i < 6; ++i) init[i] = nan;
        if (lane == 0 && a.pn >= 4) {
            int idx[4];
            p3p_select4(w, a.pn, idx);
            double X4[4][3], x4[4][2];
            for (int r = 0; r < 4; ++r) {
                for (int c = 0; c < 3; ++c) X4[r][c] = p3[3 * idx[r] + c];
                x4[r][0] = p2[2 * idx[r]]; x4[r][1] = p2[2 * idx[r] + 1];
            }
            p3p_solve4(X4, x4, cam, init);
        }
#pragma unroll
        for (int i = 0; i < 6; ++i) init[i] = __shfl_sync(0xffffffffu, init[i], 0);
    }
    if (a.init_out && lane < 6) a.init_out[(size_t)prob * 6 + lane] = init[lane];
    PnpOptions opt;
    opt.max_num_iterations = a.max_num_iterations; opt.function_tolerance = a.function_tolerance;
    opt.gradient_tolerance = a.gradient_tolerance; opt.parameter_tolerance = a.parameter_tolerance;
    PnpState st;
    PnpNormal n;
    pnp_warp_normal_at(init, p2, p3, w, cam, a.pn, lane, n);
    pnp_init(st, init, n);
    double cand[6];
    while (pnp_propose(st, opt, cand)) {
        pnp_warp_normal_at(cand, p2, p3, w, cam, a.pn, lane, n);
        if (!pnp_update(st, opt, cand, n)) break;
    }
    if (lane < 6) a.result_rt[(size_t)prob * 6 + lane] = st.x[lane];
    if (a.info && lane == 0) { a.info[2 * prob] = st.iterations; a.info[2 * prob + 1] = st.code; }
}

cudaError_t launch_pnp_fused(const PnpFusedArgs &a, cudaStream_t st)
{
    if (a.n <= 0) return cudaSuccess;
    pnp_fused_kernel<<<(a.n + PNPF_WARPS - 1) / PNPF_WARPS, PNPF_WARPS * 32, 0, st>>>(a);
    return cudaGetLastError();
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
