// exchange.cu -- multi-GPU result exchange over NVLink peer memory.
//
// The voting layer shards by image and has NO exchange step inside the algorithm; the only cross-GPU traffic is the
// [B/G, K, 2] keypoints of every rank becoming visible on every rank.  The reference has nothing here (DataParallel in
// the trainer only, lib/train/trainers/trainer.py:5,11).  Round 1 used one NCCL all_gather per call: at 8 GPUs the NCCL
// kernel (co-scheduled with the vote kernel, spinning until the slowest rank arrives) cost 0.30 ms of a 0.97 ms step.
//
// Here the exchange is part of the producing kernel: every rank owns a receive ring in its own HBM
//     recv  [slots][world][bytes_per_rank]      flags uint64 [slots][world]
// mapped into every peer (CUDA IPC, opened once).  The refit kernel's last CTA stores the rank's result block into slot
// (seq-1) % slots of EVERY peer's ring -- plain stores that travel over NVLink/NVSwitch -- and then publishes `seq` in the
// peers' flag words (vote.cu, "exchange tail").  Producers never wait.  A consumer that wants the gathered result of call
// `seq` enqueues pvb_exchange_wait on its stream: one small CTA that polls its OWN HBM until all flags[slot][r] >= seq and
// copies the slot out.  Ring reuse is made safe by the caller's schedule (clean_pvnet_b200/parallel.py: the wait of call
// s-D is enqueued before call s, slots = 2*D), not by acknowledgements, so no kernel ever blocks on a peer's progress
// except the wait kernel itself, and that one is bounded by a timeout.
#include <cstdio>
#include <cstring>
#include <new>
#include "kernels.h"

namespace pvb {

__device__ __forceinline__ unsigned long long ld_acquire_sys(const unsigned long long *p)
{
    unsigned long long v;
    asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}

__device__ __forceinline__ unsigned long long global_ns()
{
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}

constexpr int XW_THREADS = 128;

__global__ void __launch_bounds__(XW_THREADS)
exchange_wait_kernel(const unsigned long long *__restrict__ flags, unsigned long long seq, const uint4 *__restrict__ recv,
                     uint4 *__restrict__ out, size_t n16, int world, unsigned long long timeout_ns, int *status)
{
    __shared__ int s_timed_out;
    const int tid = threadIdx.x;
    if (tid == 0) s_timed_out = 0;
    __syncthreads();
    if (tid < world) {
        const unsigned long long t0 = global_ns();
        while (ld_acquire_sys(flags + tid) < seq) {
            if (global_ns() - t0 > timeout_ns) { atomicExch(&s_timed_out, 1); break; }
            __nanosleep(100);
        }
    }
    __syncthreads();
    if (s_timed_out) {
        if (tid == 0) atomicExch(status, 1);
        const uint4 nan4 = make_uint4(0x7fc00000u, 0x7fc00000u, 0x7fc00000u, 0x7fc00000u);
        for (size_t i = tid; i < n16; i += XW_THREADS) out[i] = nan4;
        return;
    }
    // the peers' stores were released before their flags; the acquire loads above order these reads after them
    for (size_t i = tid; i < n16; i += XW_THREADS) out[i] = __ldcg(recv + i);
}

cudaError_t launch_exchange_wait(const unsigned long long *flags, unsigned long long seq, const void *recv, void *out,
                                 size_t n16, int world, unsigned long long timeout_ns, int *status, cudaStream_t st)
{
    exchange_wait_kernel<<<1, XW_THREADS, 0, st>>>(flags, seq, static_cast<const uint4 *>(recv), static_cast<uint4 *>(out),
                                                   n16, world, timeout_ns, status);
    return cudaGetLastError();
}

} // namespace pvb
