// Kernel arguments of one fused pass, shared by fused_adrt.cu (fused_tile.h kernels) and
// stream_adrt.cu (stream_tile.h kernels).
#pragma once

#include "common.cuh"
#include "fused_plan.h"

namespace adrt_b200 {

struct PassArgs {
    int n, D, e, loge, next_g, d_need;
    long long in_pitch, out_pitch;
    long long src_plane_stride, dst_plane_stride;  // elements between consecutive planes
    int planes;
    int x_off;              // first d-tile of this launch (streaming passes launch interior and boundary tiles separately)
    int y_off;              // first group of this launch (angle-block sharding runs a rank's share of the groups)
    int q_first, q_count;   // image loader: plane -> (image = plane / q_count, quadrant = q_first + plane % q_count)
    int plane0;             // image loader: index of this launch's first plane in the batch (waves, see run_plan)
    int side_idx;           // host only: which helper stream (aux_stream) takes the boundary tiles of this launch
    int skip_zero;          // transposed: all-zero tiles are not written (plan::Pass::skip_zero)
    int sup_loge, sup_gmask;   // transposed: geometry of the producer that skipped them (TileCtx::sup_*)
    long long sub_delta = 0;   // transposed public-layout loader: subtract the sinogram this many bytes away (TileCtx::sub_delta)
};

// Persistent, plane-ordered scheduling of one pass (sched.cuh); next == nullptr: ordinary grid launch.
struct SchedArgs {
    unsigned *next;         // work counter of this launch, zeroed by the host
    const unsigned *dep;    // per-plane counters bumped by the producing pass (nullptr: no dependency)
    unsigned *done;         // per-plane counters this pass bumps after every tile (nullptr: nobody waits)
    unsigned dep_need;      // dep[plane] at which the plane's input is complete
    int tiles_x, tiles_y;   // tile grid of one plane (d-tile index = x + PassArgs::x_off)
    unsigned total;         // planes * tiles_x * tiles_y
    int ctas;               // persistent grid size
    int cap_per_sm;         // > 0: pad the dynamic shared memory until at most this many CTAs fit on an SM
};

// dynamic shared memory (>= base) with which at most `cap` CTAs of `kern` are resident per SM
// (cap <= 0: base).  Cached per kernel instantiation by the callers.
template <typename Kern>
inline size_t capped_smem(Kern kern, int threads, size_t base, int cap)
{
    if (cap <= 0) return base;
    size_t smem = base;
    for (;;) {
        int occ = 0;
        if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, threads, smem) != cudaSuccess) {
            (void)cudaGetLastError();
            return base;
        }
        if (occ <= cap || smem + 1024 > 226 * 1024) return smem;
        smem += 1024;
    }
}

// fp32 streaming passes (plan::Pass::stream); defined in stream_adrt.cu
// helper streams of the current device (0: boundary tiles, 1: odd waves, 2: boundary tiles of odd waves); nullptr on failure
cudaStream_t aux_stream(int idx);

int launch_stream_pass(const plan::Pass &p, bool forward, const float *src, float *dst, const PassArgs &a, cudaStream_t s);
// staged variants of the streaming passes (plan::Pass::staged); defined in stage_adrt.cu
int launch_staged_pass(const plan::Pass &p, bool forward, const float *src, float *dst, const PassArgs &a, cudaStream_t s);
// false when the driver has no tensor-map encoder: the plain streaming kernels (same tile geometry) run instead
bool staged_pass_available();
// persistent variant: `sc` describes the whole pass (all its d-tiles); the transposed direction splits it
// into an interior and a boundary launch, the latter on `side` with its own work counter sc.next + 1
int launch_stream_pass_sched(const plan::Pass &p, bool forward, const float *src, float *dst, const PassArgs &a,
                             const SchedArgs &sc, cudaStream_t s, cudaStream_t side);

}  // namespace adrt_b200
