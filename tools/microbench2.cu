// microbench2.cu -- does packed fp32 (fma.rn.f32x2, sm_100) make the cone test cheaper?
// Same block as tools/microbench.cu mode 0, but two hypotheses share one f32x2 FMA; the pixel record is stored
// duplicated ((A1,A1),(A2,A2),...) so every operand is an aligned 64-bit pair straight from LDS.128.
#include <cstdio>
#include <cuda_runtime.h>
#include <math_constants.h>

constexpr int HPT = 4, BLOCK = 16, TILE = 256;
typedef unsigned long long u64;

__device__ __forceinline__ void lds128x2(unsigned addr, u64 &a, u64 &b)
{
    asm volatile("ld.shared.v2.b64 {%0,%1}, [%2];" : "=l"(a), "=l"(b) : "r"(addr) : "memory");
}
__device__ __forceinline__ u64 fma2(u64 a, u64 b, u64 c)
{
    u64 d;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
    return d;
}
__device__ __forceinline__ u64 pack(float lo, float hi)
{
    u64 d;
    asm("mov.b64 %0, {%1,%2};" : "=l"(d) : "f"(lo), "f"(hi));
    return d;
}
__device__ __forceinline__ void unpack(u64 v, float &lo, float &hi)
{
    asm("mov.b64 {%0,%1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v));
}

template <int MODE>
__global__ void __launch_bounds__(128, 8) k(const float *in, float *out, int niter)
{
    // per pixel: 6 duplicated coefficients = 12 floats = 3 x 16 B
    __shared__ __align__(16) float s_r[TILE * 12];
    for (int i = threadIdx.x; i < TILE; i += 128)
        for (int c = 0; c < 6; ++c) { s_r[i * 12 + 2 * c] = in[i + c]; s_r[i * 12 + 2 * c + 1] = in[i + c]; }
    __syncthreads();
    u64 hx2[HPT / 2], hy2[HPT / 2];
    int neg[HPT];
    for (int j = 0; j < HPT / 2; ++j) {
        hx2[j] = pack(in[threadIdx.x + 2 * j], in[threadIdx.x + 2 * j + 1]);
        hy2[j] = pack(in[threadIdx.x + 2 * j + 7], in[threadIdx.x + 2 * j + 8]);
    }
    for (int j = 0; j < HPT; ++j) neg[j] = 0;
    const unsigned s0 = (unsigned)__cvta_generic_to_shared(s_r);
    float mnall = CUDART_INF_F;
    for (int it = 0; it < niter; ++it) {
        const int i0 = (it * BLOCK) & (TILE - 1);
        float mn[HPT];
#pragma unroll
        for (int j = 0; j < HPT; ++j) mn[j] = CUDART_INF_F;
#pragma unroll
        for (int u = 0; u < BLOCK; u += 2) {
            u64 A1a, A2a, A3a, B1a, B2a, B3a, A1b, A2b, A3b, B1b, B2b, B3b;
            const unsigned pa = s0 + (i0 + u) * 48, pb = pa + 48;
            lds128x2(pa, A1a, A2a); lds128x2(pa + 16, A3a, B1a); lds128x2(pa + 32, B2a, B3a);
            lds128x2(pb, A1b, A2b); lds128x2(pb + 16, A3b, B1b); lds128x2(pb + 32, B2b, B3b);
#pragma unroll
            for (int j = 0; j < HPT / 2; ++j) {
                const u64 ap0 = fma2(A1a, hx2[j], fma2(A2a, hy2[j], A3a));
                const u64 pp0 = fma2(B1a, hx2[j], fma2(B2a, hy2[j], B3a));
                const u64 ap1 = fma2(A1b, hx2[j], fma2(A2b, hy2[j], A3b));
                const u64 pp1 = fma2(B1b, hx2[j], fma2(B2b, hy2[j], B3b));
                float a0l, a0h, p0l, p0h, a1l, a1h, p1l, p1h;
                unpack(ap0, a0l, a0h); unpack(pp0, p0l, p0h); unpack(ap1, a1l, a1h); unpack(pp1, p1l, p1h);
                const float m00 = a0l - fabsf(p0l), m01 = a0h - fabsf(p0h);   // pixel a: hyps 2j, 2j+1
                const float m10 = a1l - fabsf(p1l), m11 = a1h - fabsf(p1h);   // pixel b
                neg[2 * j] += (int)(__float_as_uint(m00) >> 31); neg[2 * j] += (int)(__float_as_uint(m10) >> 31);
                neg[2 * j + 1] += (int)(__float_as_uint(m01) >> 31); neg[2 * j + 1] += (int)(__float_as_uint(m11) >> 31);
                if (MODE == 0) {
                    mn[2 * j] = fminf(mn[2 * j], fminf(fabsf(m00), fabsf(m10)));
                    mn[2 * j + 1] = fminf(mn[2 * j + 1], fminf(fabsf(m01), fabsf(m11)));
                }
            }
        }
#pragma unroll
        for (int j = 0; j < HPT; ++j) mnall = fminf(mnall, mn[j]);
    }
    float r = mnall;
    for (int j = 0; j < HPT; ++j) r += (float)neg[j];
    out[blockIdx.x * 128 + threadIdx.x] = r;
}

template <int MODE>
void run(const char *name, const float *in, float *out, int ctas_per_sm)
{
    const int niter = 4096;
    int sms = 148, clk = 1965000;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
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
    const double cycles = ms * 1e-3 * (double)clk * 1e3;
    printf("%-44s ctas/SM=%2d  %.3f ms  %.1f cycles per 16-px block per SMSP (4 hyps/thread)  err=%s\n", name, ctas_per_sm, ms,
           cycles / ((double)niter * ctas_per_sm), cudaGetErrorString(cudaGetLastError()));
}

int main()
{
    float *in, *out;
    cudaMalloc(&in, 4096 * sizeof(float));
    cudaMalloc(&out, 148 * 16 * 128 * sizeof(float));
    float h[4096];
    for (int i = 0; i < 4096; ++i) h[i] = 0.001f * (float)((i * 7919) % 1000) - 0.5f;
    cudaMemcpy(in, h, sizeof(h), cudaMemcpyHostToDevice);
    for (int occ : {4, 8}) {
        run<0>("f32x2: 8 FFMA2 + 4 FADD + 4 LEA + 2 FMNMX3 /px", in, out, occ);
        run<1>("f32x2: no min tracking", in, out, occ);
    }
    return 0;
}
