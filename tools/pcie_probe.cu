// pcie_probe.cu -- what does a B200 pay for reading pinned HOST memory in place, by access pattern?
//
// Question behind it (DESIGN.md "end-to-end"): the host entry fetches only the SELECTED pixels' vertex rows (72 B each, 8-byte
// aligned, about one pixel in ten) straight from the pinned tensor.  Is the link charged 72 B, 96 B (32-byte sectors) or
// 128 B (64-byte blocks) per row, and how does that compare with a DMA?
//   build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/pcie_probe.bin tools/pcie_probe.cu
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(1); } } while (0)

// every `lanes` consecutive threads read one chunk of lanes*8 bytes that starts at  base + i*stride + skew(i)
__global__ void chunk_read(const unsigned long long *__restrict__ src, size_t nchunks, int lanes, size_t stride8, int skew_mod,
                           unsigned long long *sink)
{
    const size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    const size_t total = nchunks * lanes;
    unsigned long long acc = 0;
    for (size_t e = t; e < total; e += (size_t)gridDim.x * blockDim.x) {
        const size_t i = e / lanes;
        const int l = (int)(e - i * lanes);
        const size_t off = i * stride8 + (skew_mod ? (i * 2654435761ull >> 7) % skew_mod : 0) + l;
        acc += __ldg(src + off);
    }
    if (acc == 0x1234567ull) *sink = acc;
}

__global__ void stream_read(const uint4 *__restrict__ src, size_t n16, unsigned long long *sink)
{
    unsigned long long acc = 0;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n16; i += (size_t)gridDim.x * blockDim.x) {
        const uint4 v = __ldg(src + i);
        acc += v.x + v.y + v.z + v.w;
    }
    if (acc == 0x1234567ull) *sink = acc;
}

int main()
{
    const size_t bytes = (size_t)768 << 20;
    unsigned long long *host = nullptr, *dev = nullptr, *sink = nullptr;
    CK(cudaHostAlloc((void **)&host, bytes, cudaHostAllocMapped));
    for (size_t i = 0; i < bytes / 8; ++i) host[i] = i * 3 + 1;
    CK(cudaMalloc((void **)&dev, bytes));
    CK(cudaMalloc((void **)&sink, 8));
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    float ms;
    // DMA reference
    for (int rep = 0; rep < 2; ++rep) {
        CK(cudaEventRecord(e0));
        CK(cudaMemcpyAsync(dev, host, bytes, cudaMemcpyHostToDevice));
        CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
        CK(cudaEventElapsedTime(&ms, e0, e1));
    }
    printf("DMA  cudaMemcpyAsync H2D %zu MB: %.3f ms  %.1f GB/s\n", bytes >> 20, ms, bytes / ms / 1e6);
    for (int rep = 0; rep < 2; ++rep) {
        CK(cudaEventRecord(e0));
        stream_read<<<148 * 8, 256>>>((const uint4 *)host, (bytes / 3) / 16, sink);
        CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
        CK(cudaEventElapsedTime(&ms, e0, e1));
    }
    printf("in-place contiguous 16-byte loads %zu MB: %.3f ms  %.1f GB/s\n", (bytes / 3) >> 20, ms, (bytes / 3) / ms / 1e6);
    // chunked reads: payload 24 MB each, stride 720 B (one 72-byte row in ten)
    struct Case { int lanes; int skew_mod; const char *what; };
    const Case cases[] = {
        {4, 0, "32 B chunks, 32-B aligned"}, {8, 0, "64 B chunks, 64-B aligned"}, {8, 4, "64 B chunks, 8-B aligned (skewed)"},
        {9, 0, "72 B rows, stride 720 (every start 16-B aligned: 720 = 45*16)"}, {9, 8, "72 B rows, 8-B aligned starts (the gather's pattern)"},
        {12, 0, "96 B chunks, 32-B aligned"}, {16, 0, "128 B chunks, 128-B aligned"}, {16, 8, "128 B chunks, 8-B aligned (skewed)"},
        {32, 0, "256 B chunks, 256-B aligned"},
    };
    for (const Case &c : cases) {
        const size_t stride8 = (c.lanes == 9 ? 720 : (c.lanes <= 16 ? 768 : 2304)) / 8;    // bytes/8 between chunk starts
        const size_t nchunks = (size_t)24 * 1024 * 1024 / (c.lanes * 8);
        if ((nchunks * stride8 + 64) * 8 > bytes) { printf("skip %s\n", c.what); continue; }
        for (int rep = 0; rep < 3; ++rep) {
            CK(cudaEventRecord(e0));
            chunk_read<<<148 * 8, 256>>>(host, nchunks, c.lanes, stride8, c.skew_mod, sink);
            CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
            CK(cudaEventElapsedTime(&ms, e0, e1));
        }
        const double payload = (double)nchunks * c.lanes * 8;
        printf("in-place %-62s payload %.1f MB: %.3f ms  %.1f GB/s payload\n", c.what, payload / 1e6, ms, payload / ms / 1e6);
    }
    // ---- the same access patterns on DEVICE memory: what does HBM deliver for sparse rows?  (the gather kernel's roof)
    printf("\n-- device memory (HBM3e), L2 flushed before every repetition --\n");
    unsigned char *flush = nullptr;
    CK(cudaMalloc((void **)&flush, (size_t)512 << 20));
    for (int rep = 0; rep < 2; ++rep) {
        CK(cudaEventRecord(e0));
        stream_read<<<148 * 16, 256>>>((const uint4 *)dev, bytes / 16, sink);
        CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
        CK(cudaEventElapsedTime(&ms, e0, e1));
    }
    printf("HBM contiguous 16-byte loads %zu MB: %.3f ms  %.1f GB/s\n", bytes >> 20, ms, bytes / ms / 1e6);
    const Case dcases[] = {
        {4, 0, "32 B chunks, stride 768"}, {8, 0, "64 B chunks, stride 768"}, {9, 0, "72 B rows, stride 720"},
        {9, 8, "72 B rows, 8-B aligned starts, stride 720 (the gather's pattern)"}, {16, 0, "128 B chunks, stride 768"},
        {32, 0, "256 B chunks, stride 2304"},
    };
    for (const Case &c : dcases) {
        const size_t stride8 = (c.lanes == 9 ? 720 : (c.lanes <= 16 ? 768 : 2304)) / 8;
        const size_t nchunks = (size_t)48 * 1024 * 1024 / (c.lanes * 8);
        if ((nchunks * stride8 + 64) * 8 > bytes) { printf("skip %s\n", c.what); continue; }
        for (int rep = 0; rep < 3; ++rep) {
            CK(cudaMemsetAsync(flush, rep, (size_t)512 << 20));       // evict the 126 MB L2 between repetitions
            CK(cudaEventRecord(e0));
            chunk_read<<<148 * 16, 256>>>(dev, nchunks, c.lanes, stride8, c.skew_mod, sink);
            CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
            CK(cudaEventElapsedTime(&ms, e0, e1));
        }
        const double payload = (double)nchunks * c.lanes * 8;
        printf("HBM %-66s payload %.1f MB: %.4f ms  %.0f GB/s payload  (%.2f ns/chunk)\n", c.what, payload / 1e6, ms, payload / ms / 1e6,
               ms * 1e6 / nchunks);
    }
    return 0;
}
