// microbench.cu -- what bounds the vote kernel's inner block on sm_100a?
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/microbench tools/microbench.cu && /tmp/microbench
// Each kernel runs NITER iterations of a 16-"pixel" block for 4 "hypotheses" per thread with records read
// from shared memory (broadcast), like vote_kernel.  Variants remove instruction classes to expose the
// per-pipe rates.  Reports cycles per block per SMSP-resident warp set and instructions/cycle.
#include <cstdio>
#include <cuda_runtime.h>
#include <math_constants.h>

constexpr int HPT = 4, BLOCK = 16, TILE = 256;

__device__ __forceinline__ float4 lds128(unsigned addr)
{
    float4 v;
    asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr) : "memory");
    return v;
}

template <int MODE>
__global__ void __launch_bounds__(128, 8) k(const float *in, float *out, int niter)
{
    __shared__ __align__(16) float4 s_a[TILE];
    __shared__ __align__(16) float2 s_b[TILE];
    for (int i = threadIdx.x; i < TILE; i += 128) {
        s_a[i] = make_float4(in[i], in[i + 1], in[i + 2], in[i + 3]);
        s_b[i] = make_float2(in[i + 4], in[i + 5]);
    }
    __syncthreads();
    float hxc[HPT], hyc[HPT];
    int neg[HPT];
    float acc[HPT];
    for (int j = 0; j < HPT; ++j) { hxc[j] = in[threadIdx.x + j]; hyc[j] = in[threadIdx.x + j + 7]; neg[j] = 0; acc[j] = 0.f; }
    const unsigned sa0 = (unsigned)__cvta_generic_to_shared(s_a), sb0 = (unsigned)__cvta_generic_to_shared(s_b);
    float mnall = CUDART_INF_F;
    const unsigned two = (unsigned)(in[3000] * 0.f) + 2u;   // opaque to the compiler
    for (int it = 0; it < niter; ++it) {
        const int i0 = (it * BLOCK) & (TILE - 1);
        const unsigned sa = sa0 + i0 * 16, sb = sb0 + i0 * 8;
        float mn[HPT];
#pragma unroll
        for (int j = 0; j < HPT; ++j) mn[j] = CUDART_INF_F;
#pragma unroll
        for (int u = 0; u < BLOCK; u += 2) {
            const float4 ra0 = lds128(sa + u * 16), ra1 = lds128(sa + u * 16 + 16);
            const float4 rbb = lds128(sb + u * 8);
#pragma unroll
            for (int j = 0; j < HPT; ++j) {
                if (MODE == 0 || MODE == 1 || MODE == 2) {      // full margin
                    const float ap0 = fmaf(ra0.x, hxc[j], fmaf(ra0.y, hyc[j], ra0.z));
                    const float pp0 = fmaf(ra0.w, hxc[j], fmaf(rbb.x, hyc[j], rbb.y));
                    const float ap1 = fmaf(ra1.x, hxc[j], fmaf(ra1.y, hyc[j], ra1.z));
                    const float pp1 = fmaf(ra1.w, hxc[j], fmaf(rbb.z, hyc[j], rbb.w));
                    const float m0 = ap0 - fabsf(pp0), m1 = ap1 - fabsf(pp1);
                    if (MODE == 0) {                             // everything
                        neg[j] += (int)(__float_as_uint(m0) >> 31);
                        neg[j] += (int)(__float_as_uint(m1) >> 31);
                        mn[j] = fminf(mn[j], fminf(fabsf(m0), fabsf(m1)));
                    } else if (MODE == 1) {                      // no min tracking
                        neg[j] += (int)(__float_as_uint(m0) >> 31);
                        neg[j] += (int)(__float_as_uint(m1) >> 31);
                    } else {                                     // FMA pipe only (keep results alive cheaply)
                        acc[j] += m0; acc[j] += m1;              // 2 more FADD per pair
                    }
                } else if (MODE == 3) {                          // half the FFMAs (a' only), full ALU work
                    const float ap0 = fmaf(ra0.x, hxc[j], fmaf(ra0.y, hyc[j], ra0.z));
                    const float ap1 = fmaf(ra1.x, hxc[j], fmaf(ra1.y, hyc[j], ra1.z));
                    neg[j] += (int)(__float_as_uint(ap0) >> 31);
                    neg[j] += (int)(__float_as_uint(ap1) >> 31);
                    mn[j] = fminf(mn[j], fminf(fabsf(ap0), fabsf(ap1)));
                } else if (MODE == 5 || MODE == 6 || MODE == 7) {
                    const float ap0 = fmaf(ra0.x, hxc[j], fmaf(ra0.y, hyc[j], ra0.z));
                    const float pp0 = fmaf(ra0.w, hxc[j], fmaf(rbb.x, hyc[j], rbb.y));
                    const float ap1 = fmaf(ra1.x, hxc[j], fmaf(ra1.y, hyc[j], ra1.z));
                    const float pp1 = fmaf(ra1.w, hxc[j], fmaf(rbb.z, hyc[j], rbb.w));
                    const float m0 = ap0 - fabsf(pp0), m1 = ap1 - fabsf(pp1);
                    if (MODE == 5) {            // sign tally on the FMA pipe: mad.hi.u32 (x*2)>>32 + acc
                        unsigned t0, t1;
                        asm("mad.hi.u32 %0, %1, %3, %2;" : "=r"(t0) : "r"(__float_as_uint(m0)), "r"((unsigned)neg[j]), "r"(two));
                        asm("mad.hi.u32 %0, %1, %3, %2;" : "=r"(t1) : "r"(__float_as_uint(m1)), "r"(t0), "r"(two));
                        neg[j] = (int)t1;
                        mn[j] = fminf(mn[j], fminf(fabsf(m0), fabsf(m1)));
                    } else if (MODE == 6) {     // band check by comparison (FSETP.OR) instead of FMNMX3
                        neg[j] += (int)(__float_as_uint(m0) >> 31);
                        neg[j] += (int)(__float_as_uint(m1) >> 31);
                        asm volatile("{ .reg .pred p; .reg .f32 t;\n abs.f32 t, %1; setp.lt.f32 p, t, %3;\n abs.f32 t, %2; setp.lt.or.f32 p, t, %3, p;\n @p mov.f32 %0, 0f00000000; }" : "+f"(mn[j]) : "f"(m0), "f"(m1), "f"(acc[j]));
                    } else {                    // 2-input min per test
                        neg[j] += (int)(__float_as_uint(m0) >> 31);
                        neg[j] += (int)(__float_as_uint(m1) >> 31);
                        mn[j] = fminf(mn[j], fabsf(m0));
                        mn[j] = fminf(mn[j], fabsf(m1));
                    }
                } else if (MODE == 4) {                          // two linear forms + min (FMNMX instead of FADD)
                    const float l0 = fmaf(ra0.x, hxc[j], fmaf(ra0.y, hyc[j], ra0.z));
                    const float l1 = fmaf(ra0.w, hxc[j], fmaf(rbb.x, hyc[j], rbb.y));
                    const float l2 = fmaf(ra1.x, hxc[j], fmaf(ra1.y, hyc[j], ra1.z));
                    const float l3 = fmaf(ra1.w, hxc[j], fmaf(rbb.z, hyc[j], rbb.w));
                    const float m0 = fminf(l0, l1), m1 = fminf(l2, l3);
                    neg[j] += (int)(__float_as_uint(m0) >> 31);
                    neg[j] += (int)(__float_as_uint(m1) >> 31);
                    mn[j] = fminf(mn[j], fminf(fabsf(m0), fabsf(m1)));
                }
            }
        }
#pragma unroll
        for (int j = 0; j < HPT; ++j) mnall = fminf(mnall, mn[j]);
    }
    float r = mnall;
    for (int j = 0; j < HPT; ++j) r += (float)neg[j] + acc[j];
    out[blockIdx.x * 128 + threadIdx.x] = r;
}

template <int MODE>
void run(const char *name, const float *in, float *out, int ctas_per_sm)
{
    const int niter = 4096;
    int sms = 148;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    int clk = 1965000;
    cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
    const int grid = sms * ctas_per_sm;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    k<MODE><<<grid, 128>>>(in, out, 64);
    cudaDeviceSynchronize();
    cudaEventRecord(e0);
    k<MODE><<<grid, 128>>>(in, out, niter);
    cudaEventRecord(e1);
    cudaDeviceSynchronize();
    float ms = 0;
    cudaEventElapsedTime(&ms, e0, e1);
    // warp-blocks per SMSP: ctas_per_sm CTAs x 4 warps / 4 SMSPs = ctas_per_sm warps per SMSP
    const double cycles = ms * 1e-3 * (double)clk * 1e3;
    const double per_block = cycles / ((double)niter * ctas_per_sm);
    printf("%-44s ctas/SM=%2d  %.3f ms  %.1f cycles per 16-px block per SMSP (4 hyps/thread)\n", name, ctas_per_sm, ms, per_block);
}

int main()
{
    float *in, *out;
    cudaMalloc(&in, 4096 * sizeof(float));
    cudaMalloc(&out, 148 * 16 * 128 * sizeof(float));
    float h[4096];
    for (int i = 0; i < 4096; ++i) h[i] = 0.001f * (float)((i * 7919) % 1000) - 0.5f;
    cudaMemcpy(in, h, sizeof(h), cudaMemcpyHostToDevice);
    for (int occ : {8}) {
        run<0>("0 full (16 FFMA+4 FADD+4 LEA+2 FMNMX3)/px", in, out, occ);
        run<1>("1 no min tracking", in, out, occ);
        run<2>("2 FMA pipe only (FFMA+FADD)", in, out, occ);
        run<3>("3 half FFMA, full ALU", in, out, occ);
        run<4>("4 FMNMX instead of FADD", in, out, occ);
        run<5>("5 sign tally via mad.hi.u32 (FMA pipe?)", in, out, occ);
        run<6>("6 band check via FSETP.OR per test", in, out, occ);
        run<7>("7 band check via 2-input FMNMX per test", in, out, occ);
    }
    return 0;
}
