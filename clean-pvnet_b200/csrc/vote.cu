// vote.cu -- hypothesis generation, inlier counting (the hot kernel), winner refit, covariance.
//
// Reference being replaced, per image and per round (ransac_voting_gpu.py:150-196):
//   generate_hypothesis -> zeros u8[hn,vn,tn] -> voting_for_hypothesis -> torch.sum -> torch.max
//   -> voting_for_hypothesis(hn=1) -> matmul/sum/solve
// Here: one launch each for the whole batch, no [hn,vn,tn] byte tensor, no host sync.
//
// vote_kernel design (FP32-issue bound, not HBM bound: hn tests per 16 loaded bytes):
//   * a CTA owns one (image b, keypoint k, hypothesis slice, pixel chunk);
//   * every thread keeps HPT hypotheses and their counters in registers;
//   * pixels are staged through shared memory as 6-float "cone records" and broadcast to all
//     threads (LDS.128 + LDS.64 per pixel per warp);
//   * the inlier test  cos(angle(v, h-c)) > t  is evaluated in the rotated frame of the pixel's
//     unit vector u:   a = u.(h-c),  p = u_perp.(h-c),   inlier <=> kappa*a - |p| > 0,
//     kappa = tan(acos t).  With the record (A1,A2,A3,B1,B2,B3) this is 4 FFMA + 1 FADD per test;
//   * that test is algebraically, not bitwise, the reference predicate.  A guard band delta
//     (DESIGN.md "Guard band") bounds every rounding difference between the two; whenever
//     |kappa*a-|p|| < delta the pixel is re-evaluated with the reference's exact operation
//     sequence (vote_exact).  Counts are therefore identical to the reference's.
#include <math_constants.h>
#include "common.cuh"
#include "kernels.h"

namespace pvb {

// ---------------------------------------------------------------------------------
// hypotheses: thread per (b,k,h)
// ---------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
generate_kernel(VoteArgs a)
{
    const int h = blockIdx.x * 256 + threadIdx.x;
    if (h >= a.hn) return;
    const int k = blockIdx.y, b = blockIdx.z;
    const int tn = a.tn[b];
    float x = 0.f, y = 0.f;
    if (tn > 0) {
        int t0, t1;
        if (a.idxs) {
            const int2 t = __ldg(reinterpret_cast<const int2 *>(a.idxs) + ((size_t)b * a.hn + h) * a.K + k);
            t0 = t.x; t1 = t.y;
        } else {
            const uint4 r = philox4x32_10(make_uint4((uint32_t)h, (uint32_t)k, (uint32_t)(a.img_base + b), a.tag_idx),
                                          make_uint2((uint32_t)a.seed, (uint32_t)(a.seed >> 32)));
            t0 = (int)(r.x % (uint32_t)tn);
            t1 = (int)(r.y % (uint32_t)tn);
        }
        if ((unsigned)t0 < (unsigned)tn && (unsigned)t1 < (unsigned)tn) {
            const float2 *dk = a.dirs + ((size_t)b * a.K + k) * a.cap;
            const float2 *xy = a.xy + (size_t)b * a.cap;
            const float2 d0 = dk[t0], d1 = dk[t1], c0 = xy[t0], c1 = xy[t1];
            float hx, hy;
            if (hypothesis_from_pair(d0.x, d0.y, c0.x, c0.y, d1.x, d1.y, c1.x, c1.y, hx, hy)) { x = hx; y = hy; }
        }
    }
    a.hyp[((size_t)b * a.K + k) * a.hn + h] = make_float2(x, y);
}

cudaError_t launch_generate(const VoteArgs &a, cudaStream_t st)
{
    dim3 g((a.hn + 255) / 256, a.K, a.B);
    generate_kernel<<<g, 256, 0, st>>>(a);
    return cudaGetLastError();
}

// ---------------------------------------------------------------------------------
// the vote kernel
// ---------------------------------------------------------------------------------
struct VoteK {
    VoteArgs a;
    ConeParams cone;
    int chunk;     // pixels per CTA (multiple of VOTE_TILE)
};

constexpr int VOTE_TILE = 256;

template <int HPT, int NT>
__global__ void __launch_bounds__(NT)
vote_kernel(const VoteK p)
{
    __shared__ float4 s_a[VOTE_TILE];   // (A1, A2, A3, B1)
    __shared__ float2 s_b[VOTE_TILE];   // (B2, B3)
    const VoteArgs &a = p.a;
    const int b = blockIdx.z;
    const int k = blockIdx.y % a.K, slice = blockIdx.y / a.K;
    const int tn = a.tn[b];
    const int start = blockIdx.x * p.chunk;
    if (start >= tn) return;
    const int end = min(start + p.chunk, tn);
    const int tid = threadIdx.x;

    const float ox = p.cone.ox, oy = p.cone.oy, kappa = p.cone.kappa, thresh = p.cone.thresh;
    const float cmax = a.cmax_dev ? __ldg(a.cmax_dev) : p.cone.cmax;
    const float2 *hyp = a.hyp + ((size_t)b * a.K + k) * a.hn;

    float hx[HPT], hy[HPT], hxc[HPT], hyc[HPT];
    int cnt[HPT];
    float dmax = 1e-30f;
#pragma unroll
    for (int j = 0; j < HPT; ++j) {
        const int h = slice * (NT * HPT) + j * NT + tid;
        const float2 q = (h < a.hn) ? hyp[h] : make_float2(0.f, 0.f);
        hx[j] = q.x; hy[j] = q.y;
        float xc = q.x - ox, yc = q.y - oy;
        const float S = fabsf(xc) + fabsf(yc) + cmax;
        float d = p.cone.band * S;
        if (!(S <= 1e15f) || !(d < CUDART_INF_F)) { xc = 0.f; yc = 0.f; d = CUDART_INF_F; }   // exact path only
        hxc[j] = xc; hyc[j] = yc;
        dmax = fmaxf(dmax, d);
        cnt[j] = 0;
    }

    const float2 *xy = a.xy + (size_t)b * a.cap;
    const float2 *dk = a.dirs + ((size_t)b * a.K + k) * a.cap;

    for (int t0 = start; t0 < end; t0 += VOTE_TILE) {
        const int n = min(VOTE_TILE, end - t0);
        __syncthreads();
        for (int i = tid; i < n; i += NT) {
            const float2 v = __ldg(dk + t0 + i);
            const float2 c = __ldg(xy + t0 + i);
            const float n1 = __fsqrt_rn(__fmaf_rn(v.x, v.x, __fmul_rn(v.y, v.y)));   // the reference's norm1
            float4 ra; float2 rb;
            const float cxc = c.x - ox, cyc = c.y - oy;
            if (!(n1 > __int_as_float(0x358637BD))) {
                // (double)norm1 < 1e-6 or NaN: the reference never votes for this pixel (.cu:121)
                ra = make_float4(0.f, 0.f, -1e30f, 0.f); rb = make_float2(0.f, 0.f);
            } else if (!(n1 < 1e18f) || !(fabsf(cxc) + fabsf(cyc) <= cmax)) {
                // outside the domain of the error analysis: force the exact path (m == 0 < dmax)
                ra = make_float4(0.f, 0.f, 0.f, 0.f); rb = make_float2(0.f, 0.f);
            } else {
                const float inv = 1.0f / n1;
                const float ux = v.x * inv, uy = v.y * inv;
                const float a1 = kappa * ux, a2 = kappa * uy;
                ra.x = a1; ra.y = a2; ra.z = -fmaf(a1, cxc, a2 * cyc);
                ra.w = -uy; rb.x = ux; rb.y = fmaf(uy, cxc, -(ux * cyc));
            }
            s_a[i] = ra; s_b[i] = rb;
        }
        __syncthreads();
#pragma unroll 2
        for (int i = 0; i < n; ++i) {
            const float4 ra = s_a[i];
            const float2 rb = s_b[i];
            bool f[HPT];
            float mn = CUDART_INF_F;
#pragma unroll
            for (int j = 0; j < HPT; ++j) {
                const float ap = fmaf(ra.x, hxc[j], fmaf(ra.y, hyc[j], ra.z));
                const float pp = fmaf(ra.w, hxc[j], fmaf(rb.x, hyc[j], rb.y));
                const float m = ap - fabsf(pp);
                f[j] = m > 0.f;
                mn = fminf(mn, fabsf(m));
            }
            if (__builtin_expect(mn < dmax, 0)) {
                const float2 v = __ldg(dk + t0 + i);
                const float2 c = __ldg(xy + t0 + i);
#pragma unroll
                for (int j = 0; j < HPT; ++j) f[j] = vote_exact(v.x, v.y, c.x, c.y, hx[j], hy[j], thresh);
            }
#pragma unroll
            for (int j = 0; j < HPT; ++j) cnt[j] += f[j] ? 1 : 0;
        }
    }
    int *counts = a.counts + ((size_t)b * a.K + k) * a.hn;
#pragma unroll
    for (int j = 0; j < HPT; ++j) {
        const int h = slice * (NT * HPT) + j * NT + tid;
        if (h < a.hn && cnt[j]) atomicAdd(counts + h, cnt[j]);
    }
}

// Host side of the guard band (DESIGN.md "Guard band"): u = 2^-24,
//   band = SAFETY * u * (20 + 22*kappa + 10*G),  G = 1/(t*sqrt(1-t^2)).
ConeParams make_cone(float thresh, int W, int H, float ox, float oy, bool default_origin)
{
    ConeParams c;
    c.thresh = thresh;
    if (default_origin) { ox = 0.5f * (float)(W - 1); oy = 0.5f * (float)(H - 1); }
    c.ox = ox; c.oy = oy;
    c.cmax = 0.5f * (float)(W - 1) + 0.5f * (float)(H - 1) + 1.0f;
    const double t = (double)thresh;
    if (t > 0.0 && t < 1.0) {
        const double s = sqrt(1.0 - t * t);
        const double kappa = s / t, G = 1.0 / (t * s);
        const double band = 2.0 * ldexp(1.0, -24) * (20.0 + 22.0 * kappa + 10.0 * G);
        c.kappa = (float)kappa;
        c.band = nextafterf((float)band, INFINITY);
    } else {
        c.kappa = 0.f;
        c.band = INFINITY;   // threshold outside (0,1): exact path for every test
    }
    return c;
}

cudaError_t launch_vote(const VoteArgs &a, cudaStream_t st)
{
    cudaError_t e = cudaMemsetAsync(a.counts, 0, sizeof(int) * (size_t)a.B * a.K * a.hn, st);
    if (e != cudaSuccess) return e;
    VoteK p;
    p.a = a;
    p.cone = make_cone(a.thresh, a.W, a.H, a.ox, a.oy, a.cmax_dev == nullptr);
    p.chunk = 1024;
    const int chunks = (a.cap + p.chunk - 1) / p.chunk;
#define PVB_VOTE(HPT, NT)                                                               \
    do {                                                                                \
        const int slices = (a.hn + (HPT) * (NT) - 1) / ((HPT) * (NT));                  \
        dim3 g(chunks, a.K * slices, a.B);                                              \
        vote_kernel<HPT, NT><<<g, NT, 0, st>>>(p);                                      \
    } while (0)
    if (a.hn <= 128) PVB_VOTE(1, 128);
    else if (a.hn <= 256) PVB_VOTE(2, 128);
    else PVB_VOTE(4, 128);
#undef PVB_VOTE
    return cudaGetLastError();
}

// ---------------------------------------------------------------------------------
// winner (torch.max semantics: first maximal index, ransac_voting_gpu.py:160-167) + least-squares
// refit over the winner's inliers (:177-196).  One CTA per (image, keypoint).
// ---------------------------------------------------------------------------------
constexpr int RF_THREADS = 256;

__global__ void __launch_bounds__(RF_THREADS)
refit_kernel(VoteArgs a, float2 *__restrict__ win, float *__restrict__ out)
{
    const int k = blockIdx.x, b = blockIdx.y;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int tn = a.tn[b];
    const size_t bk = (size_t)b * a.K + k;
    if (a.state[b] != 0 || tn <= 0) {   // :129-132 -> zeros
        if (tid == 0) { out[bk * 2] = 0.f; out[bk * 2 + 1] = 0.f; win[bk] = make_float2(0.f, 0.f); }
        return;
    }
    const int *counts = a.counts + bk * a.hn;
    int bc = -1, bh = 0x7fffffff;
    for (int h = tid; h < a.hn; h += RF_THREADS) {
        const int c = counts[h];
        if (c > bc) { bc = c; bh = h; }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const int oc = __shfl_xor_sync(0xffffffffu, bc, o), oh = __shfl_xor_sync(0xffffffffu, bh, o);
        if (oc > bc || (oc == bc && oh < bh)) { bc = oc; bh = oh; }
    }
    __shared__ int s_c[RF_THREADS / 32], s_h[RF_THREADS / 32];
    __shared__ float2 s_win;
    __shared__ double s_acc[RF_THREADS / 32][5];
    if (lane == 0) { s_c[warp] = bc; s_h[warp] = bh; }
    __syncthreads();
    if (tid == 0) {
        for (int w = 1; w < RF_THREADS / 32; ++w)
            if (s_c[w] > bc || (s_c[w] == bc && s_h[w] < bh)) { bc = s_c[w]; bh = s_h[w]; }
        // all_win_ratio starts at 0 and is replaced only by a strictly larger ratio (:165-167)
        s_win = (bc > 0) ? a.hyp[bk * a.hn + bh] : make_float2(0.f, 0.f);
        win[bk] = s_win;
    }
    __syncthreads();
    const float wx = s_win.x, wy = s_win.y;
    const float2 *xy = a.xy + (size_t)b * a.cap;
    const float2 *dk = a.dirs + bk * a.cap;
    double a00 = 0, a01 = 0, a11 = 0, b0 = 0, b1 = 0;
    for (int t = tid; t < tn; t += RF_THREADS) {
        const float2 v = __ldg(dk + t), c = __ldg(xy + t);
        if (vote_exact(v.x, v.y, c.x, c.y, wx, wy, a.thresh)) {
            const double nx = (double)v.y, ny = -(double)v.x;       // normal = (d_y, -d_x)  (:178-180)
            const double bb = nx * (double)c.x + ny * (double)c.y;   // b = n . c             (:189)
            a00 += nx * nx; a01 += nx * ny; a11 += ny * ny;          // ATA                   (:190)
            b0 += nx * bb; b1 += ny * bb;                            // ATb                   (:191)
        }
    }
    a00 = warp_sum(a00); a01 = warp_sum(a01); a11 = warp_sum(a11); b0 = warp_sum(b0); b1 = warp_sum(b1);
    if (lane == 0) { s_acc[warp][0] = a00; s_acc[warp][1] = a01; s_acc[warp][2] = a11; s_acc[warp][3] = b0; s_acc[warp][4] = b1; }
    __syncthreads();
    if (tid == 0) {
        double s[5] = {0, 0, 0, 0, 0};
        for (int w = 0; w < RF_THREADS / 32; ++w)
            for (int i = 0; i < 5; ++i) s[i] += s_acc[w][i];
        const double det = s[0] * s[2] - s[1] * s[1];
        float x, y;
        if (det == 0.0 || !isfinite(det)) { x = (float)s[3]; y = (float)s[4]; }   // b_inv's identity fallback (:105-108)
        else { x = (float)((s[2] * s[3] - s[1] * s[4]) / det); y = (float)((s[0] * s[4] - s[1] * s[3]) / det); }
        out[bk * 2] = x; out[bk * 2 + 1] = y;
    }
}

cudaError_t launch_refit(const VoteArgs &a, float2 *win, float *out_kpt, cudaStream_t st)
{
    dim3 g(a.K, a.B);
    refit_kernel<<<g, RF_THREADS, 0, st>>>(a, win, out_kpt);
    return cudaGetLastError();
}

// ---------------------------------------------------------------------------------
// covariance of the hypothesis cloud (ransac_voting_gpu.py:243-244, 254-269)
// ---------------------------------------------------------------------------------
constexpr int CV_THREADS = 256;

__global__ void __launch_bounds__(CV_THREADS)
covariance_kernel(VoteArgs a, const float *__restrict__ mean, float *__restrict__ cov)
{
    const int k = blockIdx.x, b = blockIdx.y;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const size_t bk = (size_t)b * a.K + k;
    const int tn = a.tn[b];
    const bool skipped = a.state[b] != 0;          // :211-216  hyp zeros, ratio ones
    const int *counts = a.counts + bk * a.hn;
    const float2 *hyp = a.hyp + bk * a.hn;
    const float ftn = (float)tn;
    __shared__ float s_max[CV_THREADS / 32];
    __shared__ double s_acc[CV_THREADS / 32][4];
    float mx = -CUDART_INF_F;
    bool has_nan = false;
    for (int h = tid; h < a.hn; h += CV_THREADS) {
        const float r = skipped ? 1.f : __fdiv_rn((float)counts[h], ftn);
        if (r != r) has_nan = true;
        mx = fmaxf(mx, r);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    if (lane == 0) s_max[warp] = mx;
    __syncthreads();
    mx = s_max[0];
    for (int w = 1; w < CV_THREADS / 32; ++w) mx = fmaxf(mx, s_max[w]);
    (void)has_nan;   // tn == 0 without skip cannot happen for min_num >= 1; ratios would be NaN as in the reference
    const float th = __fsub_rn(mx, 0.1f);
    const float mx_ = mean[bk * 2], my_ = mean[bk * 2 + 1];
    double s00 = 0, s01 = 0, s11 = 0, sw = 0;
    for (int h = tid; h < a.hn; h += CV_THREADS) {
        float w = skipped ? 1.f : __fdiv_rn((float)counts[h], ftn);
        if (w < th) w = 0.f;
        const float2 q = skipped ? make_float2(0.f, 0.f) : hyp[h];
        const double dx = (double)__fsub_rn(q.x, mx_), dy = (double)__fsub_rn(q.y, my_);
        s00 += dx * (dx * w); s01 += dx * (dy * w); s11 += dy * (dy * w); sw += w;
    }
    s00 = warp_sum(s00); s01 = warp_sum(s01); s11 = warp_sum(s11); sw = warp_sum(sw);
    if (lane == 0) { s_acc[warp][0] = s00; s_acc[warp][1] = s01; s_acc[warp][2] = s11; s_acc[warp][3] = sw; }
    __syncthreads();
    if (tid == 0) {
        double s[4] = {0, 0, 0, 0};
        for (int w = 0; w < CV_THREADS / 32; ++w)
            for (int i = 0; i < 4; ++i) s[i] += s_acc[w][i];
        const double den = (double)__fadd_rn((float)s[3], 1e-3f);
        float *c = cov + bk * 4;
        c[0] = (float)(s[0] / den); c[1] = (float)(s[1] / den);
        c[2] = (float)(s[1] / den); c[3] = (float)(s[2] / den);
    }
}

cudaError_t launch_covariance(const VoteArgs &a, const float *mean, float *out_cov, cudaStream_t st)
{
    dim3 g(a.K, a.B);
    covariance_kernel<<<g, CV_THREADS, 0, st>>>(a, mean, out_cov);
    return cudaGetLastError();
}

} // namespace pvb
