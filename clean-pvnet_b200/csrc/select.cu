// select.cu -- foreground selection: mask -> bitmap -> (thinning) -> ordered compaction + gather.
//
// Replaces, for the whole batch and without a host sync, the per-image torch ops of
//   ransac_voting_gpu.py:125-143 (v3)   cur_mask=.byte(); sum; uniform_ thinning; nonzero; masked_select
//   ransac_voting_gpu.py:207-227 (dist) cur_mask=(mask==1); ...
//
// HBM layout produced (see pvb_layout in include/pvnet_vote_b200.h):
//   bits    uint32[B][nwords]      1 bit per pixel, row-major
//   wordoff int32 [B][nwords]      exclusive popcount prefix  -> order-preserving compaction
//   xy      float2[B][cap]         (x,y) of the t-th selected pixel (torch.nonzero order, :140-141)
//   dirs    float2[B][K][cap]      vertex vectors of the selected pixels, keypoint-major so that a
//                                  (image,keypoint) vote CTA streams one contiguous float2 array
//
// All three kernels are HBM/latency bound: the mask is read exactly once (mask_bits), the
// bitmap (1/256 of an int64 mask) is what later passes touch, and the vertex field is read
// only at selected pixels.
#include "common.cuh"
#include "kernels.h"

namespace pvb {

template <typename T>
__device__ __forceinline__ uint32_t mask_byte(T v) { return (uint32_t)(uint8_t)v; }
template <>
__device__ __forceinline__ uint32_t mask_byte<float>(float v) { return (uint32_t)(uint8_t)(long long)v; }
template <>
__device__ __forceinline__ uint32_t mask_byte<double>(double v) { return (uint32_t)(uint8_t)(long long)v; }

constexpr int MB_WARPS = 8;

// One warp converts 1024 pixels into 32 bitmap words with coalesced loads (lane = pixel within
// word) and ballots; lane i keeps word i so the 32 words leave as one coalesced store.
template <typename T, int MODE>
__global__ void __launch_bounds__(MB_WARPS * 32)
mask_bits_kernel(const T *__restrict__ mask, long long sb, long long sy, long long sx, int H, int W,
                 int nwords, uint32_t *__restrict__ bits, unsigned long long *__restrict__ fgsum,
                 int *__restrict__ nz)
{
    const int b = blockIdx.y;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int w0 = (blockIdx.x * MB_WARPS + warp) * 32;
    if (w0 >= nwords) return;
    const int HW = H * W;
    const T *mb = mask + (long long)b * sb;
    const bool contig = (sx == 1 && sy == W);
    uint32_t myword = 0, sum = 0;
#pragma unroll 8
    for (int i = 0; i < 32; ++i) {
        const int p = (w0 + i) * 32 + lane;
        uint32_t val = 0;
        bool sel = false;
        if (p < HW) {
            long long off = p;
            if (!contig) { const int y = p / W; off = (long long)y * sy + (long long)(p - y * W) * sx; }
            const T v = __ldg(mb + off);
            if (MODE == PVB_SELECT_BYTE) { val = mask_byte<T>(v); sel = val != 0; }
            else { sel = (v == (T)1); val = sel; }
        }
        const uint32_t word = __ballot_sync(0xffffffffu, sel);
        if (lane == i) myword = word;
        sum += val;
    }
    if (w0 + lane < nwords) bits[(size_t)b * nwords + w0 + lane] = myword;
    const int s = warp_sum((int)sum);
    const int c = warp_sum(__popc(myword));
    if (lane == 0) {
        atomicAdd(fgsum + b, (unsigned long long)(unsigned)s);
        atomicAdd(nz + b, c);
    }
}

// One CTA per image: decides skip / thinning (ransac_voting_gpu.py:129-138), applies the
// Bernoulli thinning to the bitmap and writes the exclusive popcount prefix.
constexpr int SS_THREADS = 1024;

__global__ void __launch_bounds__(SS_THREADS)
select_scan_kernel(uint32_t *__restrict__ bits, int *__restrict__ wordoff,
                   const unsigned long long *__restrict__ fgsum, int *__restrict__ tn,
                   int *__restrict__ state, int *__restrict__ status,
                   const float *__restrict__ selection, int nwords, int HW, int min_num, int max_num,
                   int cap, uint2 key, uint32_t tag, int img_base)
{
    const int b = blockIdx.x;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const unsigned long long fg = fgsum[b];
    if (fg < (unsigned long long)(min_num < 0 ? 0 : min_num)) {   // :129  (uniform per CTA)
        if (tid == 0) { state[b] = 1; tn[b] = 0; }
        return;
    }
    const bool thin = fg > (unsigned long long)(max_num < 0 ? 0 : max_num);   // :135
    const float ratio = thin ? __fdiv_rn((float)max_num, (float)fg) : 0.f;     // max_num / fg.float()
    __shared__ int warp_tot[32];
    uint32_t *bb = bits + (size_t)b * nwords;
    int *wo = wordoff + (size_t)b * nwords;
    int base = 0;
    for (int w0 = 0; w0 < nwords; w0 += SS_THREADS) {
        const int w = w0 + tid;
        uint32_t word = (w < nwords) ? bb[w] : 0u;
        if (thin && word) {
            uint32_t keep = 0;
            if (selection) {
                const float *sp = selection + (size_t)b * HW + (size_t)w * 32;
                uint32_t m = word;
                while (m) {
                    const int j = __ffs(m) - 1;
                    m &= m - 1;
                    if (__ldg(sp + j) < ratio) keep |= 1u << j;
                }
            } else {
#pragma unroll
                for (int g = 0; g < 8; ++g) {
                    const uint32_t nib = (word >> (4 * g)) & 0xfu;
                    if (!nib) continue;
                    const uint4 r = philox4x32_10(make_uint4((uint32_t)w * 8u + g, 0u, (uint32_t)(img_base + b), tag), key);
                    uint32_t kb = 0;
                    kb |= (u32_to_unit(r.x) < ratio) ? 1u : 0u;
                    kb |= (u32_to_unit(r.y) < ratio) ? 2u : 0u;
                    kb |= (u32_to_unit(r.z) < ratio) ? 4u : 0u;
                    kb |= (u32_to_unit(r.w) < ratio) ? 8u : 0u;
                    keep |= (kb & nib) << (4 * g);
                }
            }
            word = keep;
            bb[w] = word;
        }
        const int c = __popc(word);
        int incl = c;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int v = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= o) incl += v;
        }
        if (lane == 31) warp_tot[warp] = incl;
        __syncthreads();
        if (warp == 0) {
            int t = warp_tot[lane];
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const int v = __shfl_up_sync(0xffffffffu, t, o);
                if (lane >= o) t += v;
            }
            warp_tot[lane] = t;   // inclusive prefix of warp totals
        }
        __syncthreads();
        const int warp_excl = warp ? warp_tot[warp - 1] : 0;
        const int total = warp_tot[31];
        if (w < nwords) wo[w] = base + warp_excl + incl - c;
        base += total;
        __syncthreads();
    }
    if (tid == 0) {
        state[b] = 0;
        if (base > cap) {
            atomicCAS(status, 0, PVB_ERR_CAPACITY);
            status[1] = b;
            base = cap;
        }
        tn[b] = base;
    }
}

// One warp per bitmap word: lane j owns pixel 32*w+j.  Writes xy[] and gathers the K vertex
// vectors of each selected pixel into the keypoint-major dirs[] array.
constexpr int GA_WARPS = 8;

__global__ void __launch_bounds__(GA_WARPS * 32)
gather_kernel(const uint32_t *__restrict__ bits, const int *__restrict__ wordoff,
              const int *__restrict__ state, const float *__restrict__ vertex,
              long long sB, long long sH, long long sW, long long sK, long long sC,
              float2 *__restrict__ xy, float2 *__restrict__ dirs, int nwords, int K, int cap, int W)
{
    const int b = blockIdx.y;
    if (state[b] != 0) return;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int w = blockIdx.x * GA_WARPS + warp;
    if (w >= nwords) return;
    const uint32_t word = bits[(size_t)b * nwords + w];
    if (!((word >> lane) & 1u)) return;
    const int t = wordoff[(size_t)b * nwords + w] + __popc(word & ((1u << lane) - 1u));
    if (t >= cap) return;
    const int p = w * 32 + lane;
    const int y = p / W, x = p - y * W;
    xy[(size_t)b * cap + t] = make_float2((float)x, (float)y);
    const float *vb = vertex + (long long)b * sB + (long long)y * sH + (long long)x * sW;
    float2 *db = dirs + (size_t)b * K * cap + t;
    if (sC == 1 && (sK & 1) == 0 && ((reinterpret_cast<uintptr_t>(vb) & 7u) == 0)) {
        for (int k = 0; k < K; ++k)
            db[(size_t)k * cap] = __ldg(reinterpret_cast<const float2 *>(vb + (long long)k * sK));
    } else {
        for (int k = 0; k < K; ++k) {
            const float *q = vb + (long long)k * sK;
            db[(size_t)k * cap] = make_float2(__ldg(q), __ldg(q + sC));
        }
    }
}

// ---------------------------------------------------------------------------------
cudaError_t launch_select(const SelectArgs &a, cudaStream_t st)
{
    const int nwords = a.nwords;
    dim3 g1((nwords + MB_WARPS * 32 - 1) / (MB_WARPS * 32), a.B);
#define PVB_MB(T)                                                                                        \
    do {                                                                                                 \
        if (a.select_mode == PVB_SELECT_BYTE)                                                            \
            mask_bits_kernel<T, PVB_SELECT_BYTE><<<g1, MB_WARPS * 32, 0, st>>>(                           \
                (const T *)a.mask, a.msb, a.msy, a.msx, a.H, a.W, nwords, a.bits, a.fgsum, a.nz);        \
        else                                                                                             \
            mask_bits_kernel<T, PVB_SELECT_EQ1><<<g1, MB_WARPS * 32, 0, st>>>(                            \
                (const T *)a.mask, a.msb, a.msy, a.msx, a.H, a.W, nwords, a.bits, a.fgsum, a.nz);        \
    } while (0)
    switch (a.mask_dtype) {
    case PVB_MASK_U8: PVB_MB(uint8_t); break;
    case PVB_MASK_I8: PVB_MB(int8_t); break;
    case PVB_MASK_I16: PVB_MB(int16_t); break;
    case PVB_MASK_I32: PVB_MB(int32_t); break;
    case PVB_MASK_I64: PVB_MB(long long); break;
    case PVB_MASK_F32: PVB_MB(float); break;
    case PVB_MASK_F64: PVB_MB(double); break;
    default: return cudaErrorInvalidValue;
    }
#undef PVB_MB
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return e;
    select_scan_kernel<<<a.B, SS_THREADS, 0, st>>>(a.bits, a.wordoff, a.fgsum, a.tn, a.state, a.status,
                                                   a.selection, nwords, a.H * a.W, a.min_num, a.max_num,
                                                   a.cap, make_uint2((uint32_t)a.seed, (uint32_t)(a.seed >> 32)),
                                                   a.tag_sel, a.img_base);
    e = cudaGetLastError();
    if (e != cudaSuccess) return e;
    dim3 g3((nwords + GA_WARPS - 1) / GA_WARPS, a.B);
    gather_kernel<<<g3, GA_WARPS * 32, 0, st>>>(a.bits, a.wordoff, a.state, a.vertex, a.vs[0], a.vs[1], a.vs[2],
                                                a.vs[3], a.vs[4], a.xy, a.dirs, nwords, a.K, a.cap, a.W);
    return cudaGetLastError();
}

} // namespace pvb
