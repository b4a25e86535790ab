// Persistent, plane-ordered tile scheduling shared by the fused pass kernels (fused_adrt.cu,
// stream_adrt.cu): "co-scheduled passes".
//
// A two-pass transform writes an R-layout workspace in pass 1 and reads it back in pass 2.  Launched
// one after the other over the whole batch, the workspace (17 MB forward / 34 MB transposed per
// 2048^2 fp32 plane, GBs per batch) makes a round trip through HBM.  Here both passes run AT THE SAME
// TIME as two persistent kernels on two streams:
//
//   * each kernel's CTAs pull tiles from a global work counter in plane-major order;
//   * pass 1 bumps a per-plane completion counter after every tile (red.release.gpu);
//   * a pass-2 CTA acquires its plane's counter before it loads the first tile of that plane, so
//     pass 2 trails pass 1 by a plane or two and finds the workspace in the 126 MB L2;
//   * the two kernels have different bottlenecks (pass 1 next to the public layout: shared-memory
//     pipe; streaming pass 2: latency of its bulk loads), so sharing every SM between them also
//     overlaps one kernel's load phases with the other's butterflies.
//
// Progress: the producer never waits for anything, and the consumer's shared-memory request is padded
// so that at most `k` consumer CTAs fit on an SM, which always leaves room for a producer CTA
// (cosched_plan in fused_adrt.cu).  So the producer always runs to completion and every consumer wait
// ends; the spin is bounded anyway and traps instead of hanging if that reasoning were ever violated.
#pragma once

#include "pass_args.h"

namespace adrt_b200 {

#ifdef __CUDACC__
__device__ __forceinline__ unsigned sched_ld_acquire(const unsigned *p)
{
    unsigned v;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];\n" : "=r"(v) : "l"(p) : "memory");
    return v;
}

// All threads of the CTA call this.  Returns once dep[plane] >= need was observed (acquire), i.e.
// every tile of the producing pass for this plane has been stored and is visible here.
__device__ __forceinline__ void sched_wait_plane(const SchedArgs &sc, int plane, int tid)
{
    if (tid == 0) {
        const unsigned *p = sc.dep + plane;
        unsigned spins = 0;
        while (sched_ld_acquire(p) < sc.dep_need) {
            __nanosleep(256);
            if (++spins > (1u << 24)) __trap();   // ~4 s: a broken schedule must fail, not hang the GPU
        }
    }
    __syncthreads();
    // the tile loaders read the workspace with bulk (async-proxy) copies: order them after the
    // generic-proxy acquire above
    asm volatile("fence.proxy.async;\n" ::: "memory");
}

// All threads call this after the tile's last global store.
__device__ __forceinline__ void sched_signal_plane(const SchedArgs &sc, int plane, int tid)
{
    __syncthreads();   // every thread's stores happen-before thread 0's release
    if (tid == 0) asm volatile("red.release.gpu.global.add.u32 [%0], 1;\n" ::"l"(sc.done + plane) : "memory");
}

// Work-item iteration of a persistent CTA.  `slot` is 2 words of shared memory.  The next item is
// claimed while the current tile runs, so the atomic's round trip to L2 is off the critical path.
struct SchedIter {
    unsigned it;
    __device__ __forceinline__ void begin(const SchedArgs &sc, unsigned *slot, int tid)
    {
        it = 0;
        if (tid == 0) slot[0] = atomicAdd(sc.next, 1u);
        __syncthreads();
    }
    // false: no work left
    __device__ __forceinline__ bool current(const SchedArgs &sc, unsigned *slot, int tid, int &plane, int &y, int &x)
    {
        const unsigned item = slot[it & 1];
        if (item >= sc.total) return false;
        if (tid == 0) slot[(it + 1) & 1] = atomicAdd(sc.next, 1u);
        const unsigned per_plane = (unsigned)(sc.tiles_x * sc.tiles_y);
        plane = (int)(item / per_plane);
        const unsigned rem = item - (unsigned)plane * per_plane;
        y = (int)(rem / (unsigned)sc.tiles_x);
        x = (int)(rem - (unsigned)y * (unsigned)sc.tiles_x);
        return true;
    }
    __device__ __forceinline__ void advance()
    {
        __syncthreads();   // slot[(it + 1) & 1] written by thread 0 is visible; slot[it & 1] may be reused
        ++it;
    }
};
#endif

}  // namespace adrt_b200
