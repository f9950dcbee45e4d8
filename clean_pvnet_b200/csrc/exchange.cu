// exchange.cu -- multi-GPU result exchange over NVLink peer memory.
//
// The voting layer shards by image and has NO exchange step inside the algorithm; the only cross-GPU traffic is the
// [B/G, K, 2] keypoints of every rank becoming visible on every rank.  The reference has nothing here (DataParallel in
// the trainer only, lib/train/trainers/trainer.py:5,11).  Round 1 used one NCCL all_gather per call: at 8 GPUs the NCCL
// kernel (co-scheduled with the vote kernel, spinning until the slowest rank arrives) cost 0.30 ms of a 0.97 ms step.
//
// Here the exchange is part of the producing kernel: every rank owns a receive ring in its own HBM
//     recv  uint2 [slots][world][floats_per_rank]        word = {float bits, seq}
// mapped into every peer (CUDA IPC, opened once).  The thread of the refit kernel that produces an (image, keypoint) result
// stores it into slot (seq-1) % slots of EVERY peer's ring as two 8-byte words -- plain stores that travel over
// NVLink/NVSwitch (vote.cu, "exchange tail").  An aligned 8-byte store is single-copy atomic, so each word validates
// itself (NCCL's LL protocol): no fence, no completion counter, no flag; producers never wait.  A consumer that wants the
// gathered result of call `seq` enqueues pvb_exchange_wait on its stream: one small CTA that polls the words of its OWN
// ring until all of them carry `seq`, and writes the floats out.  Ring reuse is made safe by the caller's schedule
// (clean_pvnet_b200/parallel.py: the wait of call s-D is enqueued before call s, slots = 2*D), not by acknowledgements,
// so no kernel ever blocks on a peer's progress except the wait kernel itself, and that one is bounded by a timeout.
// (Round 2 first built this with a completion counter, a system-scope fence and per-rank flag words: the fence on the
// kernel's tail cost 6 us per step, 15 us when every writer fenced -- profiles/r02_bench_n8_peer_fence_protocol.json.)
#include <cstdio>
#include <cstring>
#include <new>
#include "kernels.h"

namespace pvb {

__device__ __forceinline__ uint2 ld_word(const uint2 *p)
{
    uint2 v;
    asm volatile("ld.volatile.global.v2.u32 {%0, %1}, [%2];" : "=r"(v.x), "=r"(v.y) : "l"(p) : "memory");
    return v;
}

__device__ __forceinline__ unsigned long long global_ns()
{
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}

constexpr int XW_THREADS = 256;
constexpr int XW_BATCH = 4;          // words per thread loaded back to back before the first is examined

__global__ void __launch_bounds__(XW_THREADS)
exchange_wait_kernel(const uint2 *__restrict__ recv, unsigned int seq, float *__restrict__ out, int world, int stride_words,
                     ExchangeCounts counts, unsigned long long timeout_ns, int *status)
{
    __shared__ int s_timed_out;
    const int tid = threadIdx.x;
    if (tid == 0) s_timed_out = 0;
    __syncthreads();
    const unsigned long long t0 = global_ns();
    for (int r = 0; r < world; ++r) {
        const uint2 *src = recv + (size_t)r * stride_words;
        float *dst = out + (size_t)r * stride_words;
        const int n = counts.n[r];
        for (int i0 = tid; i0 < n; i0 += XW_THREADS * XW_BATCH) {
            uint2 w[XW_BATCH];
#pragma unroll
            for (int u = 0; u < XW_BATCH; ++u) {             // independent loads: one L2 round trip for the batch
                const int i = i0 + u * XW_THREADS;
                w[u] = (i < n) ? ld_word(src + i) : make_uint2(0u, seq);
            }
#pragma unroll
            for (int u = 0; u < XW_BATCH; ++u) {
                const int i = i0 + u * XW_THREADS;
                if (i >= n) continue;
                while (w[u].y != seq) {                      // not arrived yet (the word still carries an older seq)
                    if (global_ns() - t0 > timeout_ns) { atomicExch(&s_timed_out, 1); break; }
                    __nanosleep(100);
                    w[u] = ld_word(src + i);
                }
                dst[i] = __uint_as_float(w[u].x);
            }
        }
    }
    __syncthreads();
    if (s_timed_out) {
        if (tid == 0) atomicExch(status, 1);
        for (int r = 0; r < world; ++r)
            for (int i = tid; i < counts.n[r]; i += XW_THREADS) out[(size_t)r * stride_words + i] = __int_as_float(0x7fc00000);
    }
}

cudaError_t launch_exchange_wait(const uint2 *recv, unsigned int seq, float *out, int world, int stride_words,
                                 const ExchangeCounts &counts, unsigned long long timeout_ns, int *status, cudaStream_t st)
{
    exchange_wait_kernel<<<1, XW_THREADS, 0, st>>>(recv, seq, out, world, stride_words, counts, timeout_ns, status);
    return cudaGetLastError();
}

} // namespace pvb
