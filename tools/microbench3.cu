// microbench3.cu -- can the legacy tensor path (mma.sync m16n8k8, tf32 x3 split) carry the cone test?
//
// The cone margin m = (A1,A2,A3).(hx,hy,1) - |(B1,B2,B3).(hx,hy,1)| is a rank-3 bilinear form per (pixel, hypothesis).
// With every fp32 factor split into two tf32 numbers (x = xh + xl) and the hypothesis pre-scaled by a power of two s,
// one m16n8k8 MMA evaluates 16 pixels x 8 hypotheses of either dot product in 8 k-slots:
//     slot 0: A1h*hxh   1: A2h*hyh   2: A1h*hxl   3: A3h*s   4: A1l*hxh   5: A2l*hyh   6: A2h*hyl   7: A3l*s
// (the xl*xl terms, <= 2^-22 relative, are dropped).  This program measures, on the GPU:
//   1. that the fragment layout used by vote_mma.cu is the documented one (integer self-test);
//   2. the error of the MMA margin against a float64 evaluation of the same records, in units of 2^-24 * S;
//   3. cycles per 16-pixel x 64-hypothesis warp step: MMA only, and MMA + tally + guard-band minimum.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/microbench3.bin tools/microbench3.cu
#include <cstdio>
#include <cstdlib>
#include <cmath>
#include <vector>
#include <cuda_runtime.h>
#include <math_constants.h>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e_), __LINE__); return 1; } } while (0)

__device__ __forceinline__ float tf32_rna(float x)
{
    unsigned r;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
    return __uint_as_float(r);
}
__device__ __forceinline__ void mma_tf32(float (&d)[4], const float (&a)[4], float b0, float b1)
{
    asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%10,%10,%10,%10};"
                 : "=f"(d[0]), "=f"(d[1]), "=f"(d[2]), "=f"(d[3])
                 : "r"(__float_as_uint(a[0])), "r"(__float_as_uint(a[1])), "r"(__float_as_uint(a[2])), "r"(__float_as_uint(a[3])),
                   "r"(__float_as_uint(b0)), "r"(__float_as_uint(b1)), "f"(0.f));
}

// ---- 1. layout self-test: A[r][c] = r*8+c+1, B[k][n] = (k==n) -> D[r][n] = A[r][n]
__global__ void layout_test(float *out)
{
    const int lane = threadIdx.x, g = lane >> 2, t = lane & 3;
    float a[4] = { (float)(g * 8 + t + 1), (float)((g + 8) * 8 + t + 1), (float)(g * 8 + t + 5), (float)((g + 8) * 8 + t + 5) };
    const float b0 = (t == g) ? 1.f : 0.f, b1 = (t + 4 == g) ? 1.f : 0.f;
    float d[4];
    mma_tf32(d, a, b0, b1);
    out[g * 8 + 2 * t] = d[0]; out[g * 8 + 2 * t + 1] = d[1];
    out[(g + 8) * 8 + 2 * t] = d[2]; out[(g + 8) * 8 + 2 * t + 1] = d[3];
}

// ---- 2. accuracy: one warp per 16 pixels x 8 hypotheses; records as vote.cu builds them
struct Rec { float A1, A2, A3, B1, B2, B3; };
__global__ void accuracy(const Rec *rec, const float2 *hyp, const float *scale, float *m_out, float *p_out, float *q_out)
{
    const int lane = threadIdx.x, g = lane >> 2, t = lane & 3;
    const int blk = blockIdx.x;
    float pa[4], qa[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const Rec r = rec[blk * 16 + g + (i & 1) * 8];
        const int slot = t + (i >> 1) * 4;
        const float A1h = tf32_rna(r.A1), A2h = tf32_rna(r.A2), A3h = tf32_rna(r.A3);
        const float B1h = tf32_rna(r.B1), B2h = tf32_rna(r.B2), B3h = tf32_rna(r.B3);
        const float P[8] = { A1h, A2h, A1h, A3h, tf32_rna(r.A1 - A1h), tf32_rna(r.A2 - A2h), A2h, tf32_rna(r.A3 - A3h) };
        const float Q[8] = { B1h, B2h, B1h, B3h, tf32_rna(r.B1 - B1h), tf32_rna(r.B2 - B2h), B2h, tf32_rna(r.B3 - B3h) };
        pa[i] = P[slot]; qa[i] = Q[slot];
    }
    const float2 h = hyp[blk * 8 + g];
    const float s = scale[blk * 8 + g];
    const float hx = h.x * s, hy = h.y * s;
    const float hxh = tf32_rna(hx), hyh = tf32_rna(hy);
    const float hxl = tf32_rna(hx - hxh), hyl = tf32_rna(hy - hyh);
    const float b0 = t == 0 ? hxh : t == 1 ? hyh : t == 2 ? hxl : s;
    const float b1 = t == 0 ? hxh : t == 1 ? hyh : t == 2 ? hyl : s;
    float cp[4], cq[4];
    mma_tf32(cp, pa, b0, b1);
    mma_tf32(cq, qa, b0, b1);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int row = g + (i >> 1) * 8, col = 2 * t + (i & 1);
        const size_t o = ((size_t)blk * 16 + row) * 8 + col;
        m_out[o] = cp[i] - fabsf(cq[i]); p_out[o] = cp[i]; q_out[o] = cq[i];
    }
}

// ---- 3. throughput
template <int MODE, int NTW>
__global__ void __launch_bounds__(256, 2) thr(const float *in, int *out, int niter, long long *cyc)
{
    __shared__ __align__(16) float4 s_p[32 * 32], s_q[32 * 32];   // 32 blocks of 16 pixels = 512-pixel tile, 32 KB
    for (int i = threadIdx.x; i < 32 * 32; i += 256) {
        s_p[i] = make_float4(in[i & 1023], in[(i + 1) & 1023], in[(i + 2) & 1023], in[(i + 3) & 1023]);
        s_q[i] = make_float4(in[(i + 4) & 1023], in[(i + 5) & 1023], in[(i + 6) & 1023], in[(i + 7) & 1023]);
    }
    __syncthreads();
    const int lane = threadIdx.x & 31;
    float b0[NTW], b1[NTW];
    int neg0[NTW], neg1[NTW];
#pragma unroll
    for (int j = 0; j < NTW; ++j) { b0[j] = in[(threadIdx.x + j * 7) & 1023]; b1[j] = in[(threadIdx.x + j * 11 + 3) & 1023]; neg0[j] = neg1[j] = 0; }
    const float band = in[5] * 1e-12f;
    int flagged = 0;
    const long long c0 = clock64();
    for (int it = 0; it < niter; ++it) {
        const int blk = it & 31;
        const float4 p4 = s_p[blk * 32 + lane], q4 = s_q[blk * 32 + lane];
        const float pa[4] = { p4.x, p4.y, p4.z, p4.w }, qa[4] = { q4.x, q4.y, q4.z, q4.w };
        float mn = CUDART_INF_F;
#pragma unroll
        for (int j = 0; j < NTW; ++j) {
            float cp[4], cq[4];
            mma_tf32(cp, pa, b0[j], b1[j]);
            mma_tf32(cq, qa, b0[j], b1[j]);
            if (MODE == 0) {            // MMA only: fold the results with the cheapest possible consumer
                neg0[j] += __float_as_int(cp[0]) ^ __float_as_int(cq[3]);
            } else {
                const float m0 = cp[0] - fabsf(cq[0]), m1 = cp[1] - fabsf(cq[1]);
                const float m2 = cp[2] - fabsf(cq[2]), m3 = cp[3] - fabsf(cq[3]);
                neg0[j] += (int)(__float_as_uint(m0) >> 31); neg0[j] += (int)(__float_as_uint(m2) >> 31);
                neg1[j] += (int)(__float_as_uint(m1) >> 31); neg1[j] += (int)(__float_as_uint(m3) >> 31);
                if (MODE == 1) {
                    mn = fminf(mn, fminf(fabsf(m0), fabsf(m1)));
                    mn = fminf(mn, fminf(fabsf(m2), fabsf(m3)));
                }
            }
        }
        if (MODE == 1 && __any_sync(0xffffffffu, mn < band)) flagged++;
    }
    const long long c1 = clock64();
    int acc = flagged;
#pragma unroll
    for (int j = 0; j < NTW; ++j) acc += neg0[j] + neg1[j];
    out[blockIdx.x * 256 + threadIdx.x] = acc;
    if (threadIdx.x == 0) cyc[blockIdx.x] = c1 - c0;
}

template <int MODE, int NTW>
static int run_thr(const char *name, const float *d_in, int *d_out, long long *d_cyc, int ctas_per_sm)
{
    const int niter = 4096, grid = 148 * ctas_per_sm;
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    CK(cudaFuncSetAttribute(thr<MODE, NTW>, cudaFuncAttributeMaxDynamicSharedMemorySize, 0));
    thr<MODE, NTW><<<grid, 256>>>(d_in, d_out, niter, d_cyc);
    CK(cudaEventRecord(e0));
    thr<MODE, NTW><<<grid, 256>>>(d_in, d_out, niter, d_cyc);
    CK(cudaEventRecord(e1));
    CK(cudaDeviceSynchronize());
    float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
    std::vector<long long> cyc(grid);
    CK(cudaMemcpy(cyc.data(), d_cyc, grid * sizeof(long long), cudaMemcpyDeviceToHost));
    double mean = 0; for (long long c : cyc) mean += (double)c; mean /= grid;
    const double tests = (double)grid * 8 * niter * NTW * 128.0;
    printf("%-28s NTW=%2d ctas/SM=%d  %8.3f ms  %7.1f cyc/warp-step  %6.2f cyc per 128 tests per warp  %.2f Ttests/s\n",
           name, NTW, ctas_per_sm, ms, mean / niter, mean / niter / NTW, tests / ms * 1e-9);
    return 0;
}

int main()
{
    // 1. layout
    float *d_l; CK(cudaMalloc(&d_l, 128 * sizeof(float)));
    layout_test<<<1, 32>>>(d_l);
    float l[128]; CK(cudaMemcpy(l, d_l, sizeof(l), cudaMemcpyDeviceToHost));
    int bad = 0;
    for (int r = 0; r < 16; ++r) for (int n = 0; n < 8; ++n) if (l[r * 8 + n] != (float)(r * 8 + n + 1)) ++bad;
    printf("layout self-test: %s (%d mismatches)\n", bad ? "FAIL" : "ok", bad);

    // 2. accuracy: tile-like geometry (pixels within +-260 of the origin in L1, hypotheses up to 1e5 away), thresholds .99/.999
    for (double thr_ : { 0.99, 0.999, 0.9 }) {
        const double kappa = sqrt(1 - thr_ * thr_) / thr_;
        const int NB = 1 << 16;
        std::vector<Rec> rec(NB * 16); std::vector<float2> hyp(NB * 8); std::vector<float> sc(NB * 8);
        std::vector<double> S(NB * 8);
        srand(12345);
        auto U = []() { return (rand() + 0.5) / (RAND_MAX + 1.0); };
        std::vector<float> cx(NB * 16), cy(NB * 16), ux(NB * 16), uy(NB * 16);
        for (int b = 0; b < NB; ++b) {
            const double ext = 4 + 256 * U();            // tile half extent
            const float cmax = (float)(2 * ext) * 1.000001f + 1e-3f;
            for (int i = 0; i < 16; ++i) {
                const int o = b * 16 + i;
                cx[o] = (float)floor((2 * U() - 1) * ext) + (rand() & 1) * 0.5f; cy[o] = (float)floor((2 * U() - 1) * ext);
                const double th = 6.283185307179586 * U();
                const float vx = (float)cos(th), vy = (float)sin(th);
                const float n1 = sqrtf(fmaf(vx, vx, vy * vy)), inv = 1.0f / n1;
                ux[o] = vx * inv; uy[o] = vy * inv;
                const float a1 = (float)kappa * ux[o], a2 = (float)kappa * uy[o];
                rec[o] = { a1, a2, -fmaf(a1, cx[o], a2 * cy[o]), -uy[o], ux[o], fmaf(uy[o], cx[o], -(ux[o] * cy[o])) };
            }
            for (int j = 0; j < 8; ++j) {
                // hypotheses close to the cone boundary of pixel j (the dangerous ones) at a random distance
                const int o = b * 16 + j;
                const double dist = exp(log(2.0) + U() * log(5e4));
                const double ang = atan2((double)uy[o], (double)ux[o]) + (U() < 0.5 ? 1 : -1) * acos(thr_) * (1 + (U() - 0.5) * 1e-4);
                const float hx = (float)(cx[o] + dist * cos(ang)), hy = (float)(cy[o] + dist * sin(ang));
                hyp[b * 8 + j] = make_float2(hx, hy);
                const float Sf = fabsf(hx) + fabsf(hy) + cmax;
                int e; frexpf(Sf, &e);                   // Sf = f * 2^e, f in [0.5,1)  ->  s = 2^-e, Sf*s in [0.5,1)
                sc[b * 8 + j] = ldexpf(1.0f, -e);
                S[b * 8 + j] = Sf;
            }
        }
        Rec *d_rec; float2 *d_hyp; float *d_sc, *d_m, *d_p, *d_q;
        CK(cudaMalloc(&d_rec, rec.size() * sizeof(Rec))); CK(cudaMalloc(&d_hyp, hyp.size() * sizeof(float2)));
        CK(cudaMalloc(&d_sc, sc.size() * 4)); CK(cudaMalloc(&d_m, (size_t)NB * 128 * 4));
        CK(cudaMalloc(&d_p, (size_t)NB * 128 * 4)); CK(cudaMalloc(&d_q, (size_t)NB * 128 * 4));
        CK(cudaMemcpy(d_rec, rec.data(), rec.size() * sizeof(Rec), cudaMemcpyHostToDevice));
        CK(cudaMemcpy(d_hyp, hyp.data(), hyp.size() * sizeof(float2), cudaMemcpyHostToDevice));
        CK(cudaMemcpy(d_sc, sc.data(), sc.size() * 4, cudaMemcpyHostToDevice));
        accuracy<<<NB, 32>>>(d_rec, d_hyp, d_sc, d_m, d_p, d_q);
        CK(cudaDeviceSynchronize());
        std::vector<float> m((size_t)NB * 128), pp((size_t)NB * 128), qq((size_t)NB * 128);
        CK(cudaMemcpy(m.data(), d_m, m.size() * 4, cudaMemcpyDeviceToHost));
        CK(cudaMemcpy(pp.data(), d_p, m.size() * 4, cudaMemcpyDeviceToHost));
        CK(cudaMemcpy(qq.data(), d_q, m.size() * 4, cudaMemcpyDeviceToHost));
        // errors: (a) MMA evaluation alone = vs float64 evaluation of the SAME fp32 records and hypothesis;
        //         (b) fp32 FFMA-chain evaluation of the same (what vote.cu does today), for scale
        double emax_mma = 0, emax_ffma = 0, emax_p = 0, emax_q = 0;
        const double u = ldexp(1.0, -24);
        for (int b = 0; b < NB; ++b) for (int r = 0; r < 16; ++r) for (int n = 0; n < 8; ++n) {
            const Rec &R = rec[b * 16 + r];
            const float2 h = hyp[b * 8 + n];
            const double s = sc[b * 8 + n], Sd = S[b * 8 + n];
            const double p = (double)R.A1 * h.x + (double)R.A2 * h.y + (double)R.A3;
            const double q = (double)R.B1 * h.x + (double)R.B2 * h.y + (double)R.B3;
            const double mt = p - fabs(q);
            const size_t o = ((size_t)b * 16 + r) * 8 + n;
            const double em = fabs((double)m[o] / s - mt) / (u * Sd);
            const float pf = fmaf(R.A1, h.x, fmaf(R.A2, h.y, R.A3)), qf = fmaf(R.B1, h.x, fmaf(R.B2, h.y, R.B3));
            const double ef = fabs((double)(pf - fabsf(qf)) - mt) / (u * Sd);
            emax_mma = fmax(emax_mma, em); emax_ffma = fmax(emax_ffma, ef);
            emax_p = fmax(emax_p, fabs((double)pp[o] / s - p) / (u * Sd * kappa));
            emax_q = fmax(emax_q, fabs((double)qq[o] / s - q) / (u * Sd));
        }
        printf("accuracy thresh=%.3f kappa=%.4f : max |m_mma - m64| = %.2f u*S   (p: %.2f u*kappa*S, q: %.2f u*S)   fp32 FFMA chain: %.2f u*S\n",
               thr_, kappa, emax_mma, emax_p, emax_q, emax_ffma);
        cudaFree(d_rec); cudaFree(d_hyp); cudaFree(d_sc); cudaFree(d_m); cudaFree(d_p); cudaFree(d_q);
    }

    // 3. throughput
    float *d_in; int *d_out; long long *d_cyc;
    std::vector<float> in(1024);
    for (int i = 0; i < 1024; ++i) in[i] = (float)((rand() % 2001) - 1000) / 64.0f;
    CK(cudaMalloc(&d_in, 4096)); CK(cudaMalloc(&d_out, 148 * 4 * 256 * 4)); CK(cudaMalloc(&d_cyc, 148 * 4 * 8));
    CK(cudaMemcpy(d_in, in.data(), 4096, cudaMemcpyHostToDevice));
    for (int c = 1; c <= 2; ++c) {
        if (run_thr<0, 8>("mma only", d_in, d_out, d_cyc, c)) return 1;
        if (run_thr<2, 8>("mma + margin + tally", d_in, d_out, d_cyc, c)) return 1;
        if (run_thr<1, 8>("mma + margin + tally + min", d_in, d_out, d_cyc, c)) return 1;
        if (run_thr<1, 16>("mma + margin + tally + min", d_in, d_out, d_cyc, c)) return 1;
        if (run_thr<1, 4>("mma + margin + tally + min", d_in, d_out, d_cyc, c)) return 1;
    }
    return 0;
}
