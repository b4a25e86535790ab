// Fused multi-stage ADRT / bdrt passes for sm_100a.
//
// One CTA = one tile of one group of one (image, quadrant) plane; the tile
// algorithms are in fused_tile.h, the pass planning in fused_plan.h.  HBM
// traffic per transform drops from 2*K sinogram sweeps (per-stage kernels) to
// one sweep per pass (2 passes up to n = 4096 in fp32).
#include "pass_args.h"
#include "sched.cuh"

#include <type_traits>

namespace adrt_b200 {

namespace {


// The phases of a full tile, unrolled at compile time with a barrier after each.
template <typename Prog, typename T, int PH>
__device__ __forceinline__ void all_phases(int mode, T *buf, T (&regs)[tile::NREG], const T *sp, T *dp,
                                           const tile::TileCtx &c, int tid)
{
    if constexpr (PH < Prog::kPhases) {
        Prog::template phase_ct<PH>(mode, buf, regs, sp, dp, c, tid);
        __syncthreads();
        all_phases<Prog, T, PH + 1>(mode, buf, regs, sp, dp, c, tid);
    }
}

// resident CTAs the register allocation aims at (launch bound); ADRT_B5P_CTAS: A/B tuning of the
// fp32 five-stage transposed pass that reads the public layout (fused loader + radix-4 step)
#ifndef ADRT_B5P_CTAS
#define ADRT_B5P_CTAS 5
#endif
// ADRT_F64_F5_CTAS / ADRT_F64_B5_CTAS: same for the fp64 five-stage passes (70 KB tiles: shared
// memory allows 3).  Measured at 64 x 2048^2: 3 CTAs (80 registers) take the forward pass pair from
// 15.57 to 14.46 ms, but the transposed pair from 20.63 to 21.49 ms (its kernels spill)
#ifndef ADRT_F5I_CTAS
#define ADRT_F5I_CTAS 5
#endif
#ifndef ADRT_F64_F5_CTAS
#define ADRT_F64_F5_CTAS 3
#endif
#ifndef ADRT_F64_B5_CTAS
#define ADRT_F64_B5_CTAS 2
#endif
template <typename T, int M, int LOADK, bool kForward>
constexpr int min_ctas()
{
    if (sizeof(T) == 8 && M == 5) return kForward ? ADRT_F64_F5_CTAS : ADRT_F64_B5_CTAS;
    if (sizeof(T) == 8) return tile::Geo<M>::MIN_CTAS / 2;
    if (!kForward && LOADK == tile::LOAD_QCOLS && M == 5) return ADRT_B5P_CTAS;
    if (kForward && LOADK == tile::LOAD_IMAGE && M == 5) return ADRT_F5I_CTAS;
    return tile::Geo<M>::MIN_CTAS;
}

template <typename T, int M, int LOADK, int STOREK, bool kForward, bool kSub = false>
__global__ void __launch_bounds__(tile::Geo<M>::NT, min_ctas<T, M, LOADK, kForward>())
pass_kernel(const T *__restrict__ src, T *__restrict__ dst, PassArgs a)
{
    using Prog = typename std::conditional<kForward, tile::FwdProgram<T, M, LOADK, STOREK>,
                                           tile::BwdProgram<T, M, LOADK, STOREK, kSub>>::type;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    T *buf = reinterpret_cast<T *>(smem_raw);
    T regs[tile::NREG];

    tile::TileCtx c;
    c.n = a.n;
    c.D = a.D;
    c.e = a.e;
    c.g = blockIdx.y + a.y_off;
    c.k0 = c.g >> a.loge;
    c.a_g = c.g & (a.e - 1);
    c.d0 = blockIdx.x * Prog::TD;
    c.next_g = a.next_g;
    c.d_need = a.d_need;
    c.sup_loge = a.sup_loge;
    c.sup_gmask = a.sup_gmask;
    c.in_pitch = a.in_pitch;
    c.out_pitch = a.out_pitch;
    c.q = 0;
    if (kSub) c.sub_delta = a.sub_delta;
    const int mode = Prog::classify(c);
    if (mode == tile::TILE_SKIP || (mode == tile::TILE_ZERO && a.skip_zero)) return;

    for (int plane = blockIdx.z; plane < a.planes; plane += gridDim.z) {
        const T *sp;
        if (LOADK == tile::LOAD_IMAGE) {
            const int gp = plane + a.plane0;
            c.q = a.q_first + gp % a.q_count;
            sp = src + (long long)(gp / a.q_count) * a.src_plane_stride;
        } else {
            sp = src + (long long)plane * a.src_plane_stride;
        }
        T *dp = dst + (long long)plane * a.dst_plane_stride;
        if (mode == tile::TILE_ZERO) {
            Prog::zero_tile(buf, dp, c, threadIdx.x);
        } else {
            all_phases<Prog, T, 0>(mode, buf, regs, sp, dp, c, threadIdx.x);
        }
    }
}

template <typename T, int M, int LOADK, int STOREK, bool kForward, bool kSub = false>
int launch_pass(const T *src, T *dst, const PassArgs &a, int grid_x, int grid_y, cudaStream_t s)
{
    auto kern = pass_kernel<T, M, LOADK, STOREK, kForward, kSub>;
    size_t smem = (size_t)tile::Geo<M>::G * tile::Pitch<T>::value * sizeof(T);
    if (const char *e = getenv("ADRT_B200_SMEM_PAD_KB")) smem += (size_t)atoi(e) * 1024;  // occupancy experiments
    // per device, so not cached in a static: a process may drive several GPUs
    ADRT_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    dim3 grid((unsigned)grid_x, (unsigned)grid_y, (unsigned)(a.planes < 65535 ? a.planes : 65535));
    kern<<<grid, tile::Geo<M>::NT, smem, s>>>(src, dst, a);
    ADRT_LAUNCH_CHECK();
    return ADRT_B200_OK;
}

// Persistent variant (sched.cuh): the CTAs pull (plane, group, d-tile) work items from a global counter
// in plane-major order, wait for the producing pass of their plane where there is one and announce
// every finished tile where a later pass waits for it.
// Launch bound of the persistent kernels: they share every SM with another pass's CTAs (2-3 of them
// resident), so the register cap that buys a 4th or 5th resident CTA only costs spills here.
#ifndef ADRT_CO_CTAS
#define ADRT_CO_CTAS 3
#endif
template <typename T, int M, int LOADK, bool kForward>
constexpr int min_ctas_p()
{
    return min_ctas<T, M, LOADK, kForward>() < ADRT_CO_CTAS ? min_ctas<T, M, LOADK, kForward>() : ADRT_CO_CTAS;
}

template <typename T, int M, int LOADK, int STOREK, bool kForward>
__global__ void __launch_bounds__(tile::Geo<M>::NT, min_ctas_p<T, M, LOADK, kForward>())
pass_kernel_p(const T *__restrict__ src, T *__restrict__ dst, PassArgs a, SchedArgs sc)
{
    using Prog = typename std::conditional<kForward, tile::FwdProgram<T, M, LOADK, STOREK>,
                                           tile::BwdProgram<T, M, LOADK, STOREK>>::type;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    __shared__ unsigned slot[2];
    T *buf = reinterpret_cast<T *>(smem_raw);
    T regs[tile::NREG];
    const int tid = threadIdx.x;

    tile::TileCtx c;
    c.n = a.n;
    c.D = a.D;
    c.e = a.e;
    c.next_g = a.next_g;
    c.d_need = a.d_need;
    c.sup_loge = a.sup_loge;
    c.sup_gmask = a.sup_gmask;
    c.in_pitch = a.in_pitch;
    c.out_pitch = a.out_pitch;
    c.q = 0;
    SchedIter iter;
    iter.begin(sc, slot, tid);
    int plane, y, x, ready_plane = -1;
    while (iter.current(sc, slot, tid, plane, y, x)) {
        c.g = y + a.y_off;
        c.k0 = c.g >> a.loge;
        c.a_g = c.g & (a.e - 1);
        c.d0 = (x + a.x_off) * Prog::TD;
        const int mode = Prog::classify(c);
        if (mode != tile::TILE_SKIP && !(mode == tile::TILE_ZERO && a.skip_zero)) {
            const T *sp;
            if (LOADK == tile::LOAD_IMAGE) {
                const int gp = plane + a.plane0;
                c.q = a.q_first + gp % a.q_count;
                sp = src + (long long)(gp / a.q_count) * a.src_plane_stride;
            } else {
                sp = src + (long long)plane * a.src_plane_stride;
            }
            T *dp = dst + (long long)plane * a.dst_plane_stride;
            if (mode == tile::TILE_ZERO) {
                Prog::zero_tile(buf, dp, c, tid);
            } else {
                if (sc.dep && plane != ready_plane) {
                    sched_wait_plane(sc, plane, tid);
                    ready_plane = plane;
                }
                all_phases<Prog, T, 0>(mode, buf, regs, sp, dp, c, tid);
            }
        }
        if (sc.done) sched_signal_plane(sc, plane, tid);
        iter.advance();
    }
}

template <typename T, int M, int LOADK, int STOREK, bool kForward>
int launch_pass_p(const T *src, T *dst, const PassArgs &a, const SchedArgs &sc, cudaStream_t s)
{
    auto kern = pass_kernel_p<T, M, LOADK, STOREK, kForward>;
    ADRT_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 226 * 1024));
    static std::atomic<size_t> cached[8] = {};
    const size_t base = (size_t)tile::Geo<M>::G * tile::Pitch<T>::value * sizeof(T);
    const int ci = sc.cap_per_sm > 0 && sc.cap_per_sm < 8 ? sc.cap_per_sm : 0;
    size_t smem = cached[ci].load();
    if (!smem) cached[ci].store(smem = capped_smem(kern, tile::Geo<M>::NT, base, ci));
    kern<<<(unsigned)sc.ctas, tile::Geo<M>::NT, smem, s>>>(src, dst, a, sc);
    ADRT_LAUNCH_CHECK();
    return ADRT_B200_OK;
}

template <typename T, int M, bool kForward>
int dispatch_kinds(int load, int store, const T *src, T *dst, const PassArgs &a, int gx, int gy, cudaStream_t s)
{
    using namespace tile;
    if constexpr (kForward) {
        if (load == LOAD_IMAGE && store == STORE_WROWS) return launch_pass<T, M, LOAD_IMAGE, STORE_WROWS, kForward>(src, dst, a, gx, gy, s);
        if (load == LOAD_IMAGE && store == STORE_QCOLS) return launch_pass<T, M, LOAD_IMAGE, STORE_QCOLS, kForward>(src, dst, a, gx, gy, s);
        if (load == LOAD_WROWS && store == STORE_WROWS) return launch_pass<T, M, LOAD_WROWS, STORE_WROWS, kForward>(src, dst, a, gx, gy, s);
        if (load == LOAD_WROWS && store == STORE_QCOLS) return launch_pass<T, M, LOAD_WROWS, STORE_QCOLS, kForward>(src, dst, a, gx, gy, s);
    } else {
        if (a.sub_delta != 0) {
            // subtract-on-load: the first pass of a multi-pass transposed plan (see sub_on_load_ok)
            if constexpr (M >= 4) {
                if (load == LOAD_QCOLS && store == STORE_WROWS) return launch_pass<T, M, LOAD_QCOLS, STORE_WROWS, kForward, true>(src, dst, a, gx, gy, s);
            }
            set_error("internal: no subtract-on-load kernel for pass kinds %d/%d, %d stages", load, store, M);
            return ADRT_B200_EINVAL;
        }
        if (load == LOAD_QCOLS && store == STORE_WROWS) return launch_pass<T, M, LOAD_QCOLS, STORE_WROWS, kForward>(src, dst, a, gx, gy, s);
        if (load == LOAD_QCOLS && store == STORE_QCOLS) return launch_pass<T, M, LOAD_QCOLS, STORE_QCOLS, kForward>(src, dst, a, gx, gy, s);
        if (load == LOAD_WROWS && store == STORE_WROWS) return launch_pass<T, M, LOAD_WROWS, STORE_WROWS, kForward>(src, dst, a, gx, gy, s);
        if (load == LOAD_WROWS && store == STORE_QCOLS) return launch_pass<T, M, LOAD_WROWS, STORE_QCOLS, kForward>(src, dst, a, gx, gy, s);
    }
    set_error("internal: bad pass kinds %d/%d", load, store);
    return ADRT_B200_EINVAL;
}

template <typename T, int M, bool kForward>
int dispatch_kinds_p(int load, int store, const T *src, T *dst, const PassArgs &a, const SchedArgs &sc, cudaStream_t s)
{
    using namespace tile;
    if constexpr (kForward) {
        if (load == LOAD_IMAGE && store == STORE_WROWS) return launch_pass_p<T, M, LOAD_IMAGE, STORE_WROWS, kForward>(src, dst, a, sc, s);
        if (load == LOAD_WROWS && store == STORE_QCOLS) return launch_pass_p<T, M, LOAD_WROWS, STORE_QCOLS, kForward>(src, dst, a, sc, s);
    } else {
        if (load == LOAD_QCOLS && store == STORE_WROWS) return launch_pass_p<T, M, LOAD_QCOLS, STORE_WROWS, kForward>(src, dst, a, sc, s);
        if (load == LOAD_WROWS && store == STORE_QCOLS) return launch_pass_p<T, M, LOAD_WROWS, STORE_QCOLS, kForward>(src, dst, a, sc, s);
    }
    set_error("internal: bad co-scheduled pass kinds %d/%d", load, store);
    return ADRT_B200_EINVAL;
}

// co-scheduled (persistent) launch of one pass of a two-pass plan; `side`: helper stream for the boundary
// tiles of a streaming transposed pass
template <typename T, bool kForward>
int dispatch_pass_p(const plan::Pass &p, const T *src, T *dst, const PassArgs &a, const SchedArgs &sc, cudaStream_t s,
                    cudaStream_t side)
{
    if (p.stream) {
        if constexpr (std::is_same<T, float>::value) return launch_stream_pass_sched(p, kForward, src, dst, a, sc, s, side);
        set_error("internal: streaming passes are fp32 only");
        return ADRT_B200_EINVAL;
    }
    switch (p.M) {
    case 4: return dispatch_kinds_p<T, 4, kForward>(p.load, p.store, src, dst, a, sc, s);
    case 5: return dispatch_kinds_p<T, 5, kForward>(p.load, p.store, src, dst, a, sc, s);
    case 6: return dispatch_kinds_p<T, 6, kForward>(p.load, p.store, src, dst, a, sc, s);
    }
    set_error("internal: no co-scheduled kernel for %d stages per pass", p.M);
    return ADRT_B200_EINVAL;
}

template <typename T, bool kForward>
int dispatch_pass(const plan::Pass &p, const T *src, T *dst, const PassArgs &a, cudaStream_t s)
{
    if (p.stream) {
        if constexpr (std::is_same<T, float>::value)
            return p.staged && staged_pass_available() ? launch_staged_pass(p, kForward, src, dst, a, s)
                                                       : launch_stream_pass(p, kForward, src, dst, a, s);
        set_error("internal: streaming passes are fp32 only");
        return ADRT_B200_EINVAL;
    }
    switch (p.M) {
    case 1: return dispatch_kinds<T, 1, kForward>(p.load, p.store, src, dst, a, p.grid_x, p.grid_y, s);
    case 2: return dispatch_kinds<T, 2, kForward>(p.load, p.store, src, dst, a, p.grid_x, p.grid_y, s);
    case 3: return dispatch_kinds<T, 3, kForward>(p.load, p.store, src, dst, a, p.grid_x, p.grid_y, s);
    case 4: return dispatch_kinds<T, 4, kForward>(p.load, p.store, src, dst, a, p.grid_x, p.grid_y, s);
    case 5: return dispatch_kinds<T, 5, kForward>(p.load, p.store, src, dst, a, p.grid_x, p.grid_y, s);
    case 6: return dispatch_kinds<T, 6, kForward>(p.load, p.store, src, dst, a, p.grid_x, p.grid_y, s);
    }
    set_error("internal: unsupported stages per pass %d", p.M);
    return ADRT_B200_EINVAL;
}

// Waves: the planes of a batch are processed in waves of `planes` planes, every pass of a
// wave before the next wave starts, so that the R-layout workspace a pass writes is still
// in the 126 MB L2 when the next pass reads it (one fp32 2048^2 plane: 17 MB forward, 34 MB
// transposed).  With two lanes, odd waves run on a helper stream with their own workspace:
// the tail of one wave's kernel overlaps the head of the other's.
// ADRT_B200_WAVE_PLANES / ADRT_B200_WAVE (images) / ADRT_B200_WAVE_LANES override; 0 planes =
// the whole batch at once.
struct WaveCfg {
    int planes;  // planes per wave
    int lanes;   // 1 or 2 concurrent waves
};

// `overlap`: no override given, run the batch as two half-batch waves on two lanes (see overlap_default)
WaveCfg wave_config(int64_t total_planes, int q_count, size_t plane_ws_bytes, bool overlap)
{
    WaveCfg w;
    w.planes = (int)total_planes;
    w.lanes = 1;
    long long want = 0;
    const char *ep = getenv("ADRT_B200_WAVE_PLANES"), *ei = getenv("ADRT_B200_WAVE"), *el = getenv("ADRT_B200_WAVE_LANES");
    if (ep) want = atoll(ep);  // read per call: tunable at run time
    else if (ei) want = atoll(ei) * q_count;
    else if (overlap && !el && total_planes >= 32) {
        want = ((total_planes / 2 + q_count - 1) / q_count) * q_count;
        w.lanes = 2;
    }
    (void)plane_ws_bytes;
    if (want > 0 && want < total_planes) w.planes = (int)want;
    if (el) w.lanes = atoi(el) >= 2 ? 2 : 1;
    if ((long long)w.planes * w.lanes > total_planes) w.lanes = 1;
    return w;
}

// Forward two-pass fp32 plans whose image pass is the staged kernel run the two halves of a batch on two
// streams: the image pass (3 persistent CTAs per SM waiting on their copies most of the time, a third of the
// DRAM bandwidth) of one half overlaps the DRAM-heavy second pass of the other -- 64 x 2048^2 adrt 4.50 ->
// 4.37 ms (profiles/s6_ab_waves.jsonl).  The transposed passes lose when overlapped (6.36 -> 6.50 ms) and
// stay on one lane; smaller (L2-sized) waves lose in both directions.
template <bool kForward>
bool overlap_default(const plan::Plan &pl)
{
    return kForward && pl.npass == 2 && pl.pass[0].staged;
}

// Experiment: keep the workspace resident in L2 (persisting access-policy window on the
// streams the waves run on).  ADRT_B200_L2_PERSIST_MB = size of the persisting carve-out.
void set_persist_window(cudaStream_t s, void *base, size_t bytes)
{
    const char *e = getenv("ADRT_B200_L2_PERSIST_MB");
    if (!e) return;
    const size_t carve = (size_t)atoi(e) << 20;
    int dev = 0, max_win = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&max_win, cudaDevAttrMaxAccessPolicyWindowSize, dev);
    cudaDeviceSetLimit(cudaLimitPersistingL2CacheSize, carve);
    cudaStreamAttrValue v = {};
    v.accessPolicyWindow.base_ptr = carve ? base : nullptr;
    v.accessPolicyWindow.num_bytes = carve ? std::min(bytes, (size_t)max_win) : 0;
    v.accessPolicyWindow.hitRatio = carve ? std::min(1.0f, (float)carve / (float)std::max<size_t>(bytes, 1)) : 0.f;
    v.accessPolicyWindow.hitProp = cudaAccessPropertyPersisting;
    v.accessPolicyWindow.missProp = cudaAccessPropertyStreaming;
    cudaStreamSetAttribute(s, cudaStreamAttributeAccessPolicyWindow, &v);
    (void)cudaGetLastError();
}

// Co-scheduled two-pass plans (sched.cuh).  ADRT_B200_COSCHED=0/1 overrides the default; ADRT_B200_CO_K1 /
// ADRT_B200_CO_K2 = persistent CTAs per SM of the producing / consuming pass.
struct CoCfg {
    bool on;
    int k1, k2;
};

template <typename T>
CoCfg cosched_config(const plan::Plan &pl, int64_t total_planes)
{
    CoCfg c = {false, 2, 2};
    if (pl.npass != 2) return c;
    // Opt-in (ADRT_B200_COSCHED=1): measured slower than one launch per pass on every configuration
    // (profiles/r02_cosched_sweep.jsonl) -- every pass kernel is bound by the shared memory its
    // resident tiles occupy, so two kernels sharing an SM each run at their share of the tiles'
    // rate and the L2 hits of the consumer do not make up for it.
    c.on = false;
    if (const char *e = getenv("ADRT_B200_COSCHED")) c.on = atoi(e) != 0;
    // only the ordinary two-pass kinds (public layout on the outside) have persistent instantiations
    if (pl.pass[0].store != tile::STORE_WROWS || pl.pass[1].load != tile::LOAD_WROWS ||
        pl.pass[1].store != tile::STORE_QCOLS || pl.pass[0].load == tile::LOAD_WROWS)
        c.on = false;
    if (const char *e = getenv("ADRT_B200_CO_K1")) c.k1 = atoi(e);
    if (const char *e = getenv("ADRT_B200_CO_K2")) c.k2 = atoi(e);
    if (c.k1 < 1) c.k1 = 1;
    if (c.k2 < 1) c.k2 = 1;
    // work items are counted in 32 bits
    for (int i = 0; i < 2; ++i)
        if ((double)total_planes * pl.pass[i].grid_x * pl.pass[i].grid_y >= 4.0e9) c.on = false;
    // the masked / per-M kernels instantiated for co-scheduling
    for (int i = 0; i < 2; ++i)
        if ((!pl.pass[i].stream && (pl.pass[i].M < 4 || pl.pass[i].M > 6)) || pl.pass[i].staged) c.on = false;
    return c;
}

// counters appended to the workspace: per-plane completion counts + the work counters of the launches
inline size_t cosched_counter_words(int64_t planes) { return ((size_t)planes + 8 + 63) & ~size_t(63); }

template <typename T, bool kForward>
int run_plan_cosched(const plan::Plan &pl, const CoCfg &cc, const T *in, T *out, int64_t total, int q_first, int q_count,
                     T *ws_slot0, unsigned *counters, cudaStream_t s)
{
    const int n = pl.n, D = pl.D;
    const long long img_elems = (long long)n * n, sino_plane = (long long)D * n;
    int dev = 0, sms = 148;
    ADRT_CUDA_CHECK(cudaGetDevice(&dev));
    ADRT_CUDA_CHECK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    cudaStream_t s2 = aux_stream(1), s3 = aux_stream(3);
    if (!s2 || !s3) {
        set_error("co-scheduled passes: helper streams unavailable");
        return ADRT_B200_ECUDA;
    }
    unsigned *done = counters, *next = counters + total;   // next[0]: pass 1, next[1..2]: pass 2 (interior, boundary)
    ADRT_CUDA_CHECK(cudaMemsetAsync(counters, 0, cosched_counter_words(total) * sizeof(unsigned), s));
    cudaEvent_t fork = nullptr, join2 = nullptr, join3 = nullptr;
    ADRT_CUDA_CHECK(cudaEventCreateWithFlags(&fork, cudaEventDisableTiming));
    ADRT_CUDA_CHECK(cudaEventCreateWithFlags(&join2, cudaEventDisableTiming));
    ADRT_CUDA_CHECK(cudaEventCreateWithFlags(&join3, cudaEventDisableTiming));
    int rc = ADRT_B200_OK;
    auto fail = [&](cudaError_t e) { if (e != cudaSuccess && rc == ADRT_B200_OK) { set_error("co-scheduled passes: %s", cudaGetErrorString(e)); rc = ADRT_B200_ECUDA; } };
    fail(cudaEventRecord(fork, s));
    fail(cudaStreamWaitEvent(s2, fork, 0));
    fail(cudaStreamWaitEvent(s3, fork, 0));
    for (int i = 0; i < 2 && rc == ADRT_B200_OK; ++i) {
        const plan::Pass &p = pl.pass[i];
        PassArgs a;
        a.n = n; a.D = D; a.e = 1 << p.s; a.loge = p.s; a.next_g = p.next_g; a.d_need = p.d_need;
            a.skip_zero = p.skip_zero; a.sup_loge = p.sup_loge; a.sup_gmask = p.sup_gmask;
        a.in_pitch = p.in_pitch; a.out_pitch = p.out_pitch;
        a.planes = (int)total;
        a.x_off = 0;
        a.y_off = 0;
        a.q_first = q_first; a.q_count = q_count;
        a.plane0 = 0;
        a.side_idx = 0;
        const T *src;
        T *dst;
        if (i == 0) {
            src = in;
            a.src_plane_stride = kForward ? img_elems : (p.in_pitch ? (long long)n * p.in_pitch : sino_plane);
            dst = ws_slot0;
            a.dst_plane_stride = (long long)n * p.out_pitch;
        } else {
            src = ws_slot0;
            a.src_plane_stride = (long long)n * p.in_pitch;
            dst = out;
            a.dst_plane_stride = p.out_pitch ? (long long)n * p.out_pitch : sino_plane;
        }
        SchedArgs sc;
        sc.tiles_x = p.grid_x;
        sc.tiles_y = p.grid_y;
        sc.total = (unsigned)total * (unsigned)p.grid_x * (unsigned)p.grid_y;
        if (i == 0) {
            sc.next = next;
            sc.dep = nullptr;
            sc.done = done;
            sc.dep_need = 0;
            sc.ctas = sms * cc.k1;
            sc.cap_per_sm = 0;
            rc = dispatch_pass_p<T, kForward>(p, src, dst, a, sc, s, s);
        } else {
            sc.next = next + 1;
            sc.dep = done;
            sc.done = nullptr;
            sc.dep_need = (unsigned)pl.pass[0].grid_x * (unsigned)pl.pass[0].grid_y;
            sc.ctas = sms * cc.k2;
            sc.cap_per_sm = cc.k2;
            rc = dispatch_pass_p<T, kForward>(p, src, dst, a, sc, s2, s3);
        }
    }
    fail(cudaEventRecord(join2, s2));
    fail(cudaEventRecord(join3, s3));
    fail(cudaStreamWaitEvent(s, join2, 0));
    fail(cudaStreamWaitEvent(s, join3, 0));
    cudaEventDestroy(fork);
    cudaEventDestroy(join2);
    cudaEventDestroy(join3);
    return rc;
}

template <typename T, bool kForward>
int run_plan(const plan::Plan &pl, const T *in, T *out, int64_t B, int q_first, int q_count, T *ws, size_t ws_elems,
             cudaStream_t s, long long sub_delta = 0)
{
    // B images of q_count planes each (forward: quadrants q_first .. q_first+q_count-1 of
    // every image; transposed: any plane count, the quadrant identity does not matter)
    const int n = pl.n, D = pl.D;
    const int64_t total = B * q_count;
    const size_t plane_ws = pl.ws_slot_elems[0] + pl.ws_slot_elems[1];
    CoCfg cc = cosched_config<T>(pl, total);
    if (sub_delta != 0) cc.on = false;
    if (cc.on) {
        // whole batch, one workspace slot (pass 1 -> pass 2), counters behind it
        const size_t ws_need = pl.ws_slot_elems[0] * (size_t)total;
        const size_t ctr_elems = (cosched_counter_words(total) * sizeof(unsigned) + sizeof(T) - 1) / sizeof(T);
        const size_t ctr_off = (ws_need + 63) & ~size_t(63);
        if (ctr_off + ctr_elems > ws_elems) {
            set_error("fused workspace too small: need %zu elements, got %zu", ctr_off + ctr_elems, ws_elems);
            return ADRT_B200_EWORKSPACE;
        }
        return run_plan_cosched<T, kForward>(pl, cc, in, out, total, q_first, q_count, ws,
                                             reinterpret_cast<unsigned *>(ws + ctr_off), s);
    }
    WaveCfg wc = wave_config(total, q_count, plane_ws * sizeof(T), overlap_default<kForward>(pl));
    const size_t slot0 = pl.ws_slot_elems[0] * (size_t)wc.planes;
    const size_t lane_elems = plane_ws * (size_t)wc.planes;
    if (lane_elems * wc.lanes > ws_elems) {
        set_error("fused workspace too small: need %zu elements, got %zu", lane_elems * wc.lanes, ws_elems);
        return ADRT_B200_EWORKSPACE;
    }
    // lane 1 = helper stream, forked from / joined to the caller's stream by events
    cudaStream_t lane_stream[2] = {s, nullptr};
    cudaEvent_t fork = nullptr, join = nullptr;
    int lanes = wc.lanes;
    if (lanes == 2) {
        // the helper streams are shared per device: never fork into them while the caller's stream is being
        // captured into a graph (the waves then simply follow one another on the caller's stream)
        cudaStreamCaptureStatus cap = cudaStreamCaptureStatusNone;
        if (cudaStreamIsCapturing(s, &cap) != cudaSuccess || cap != cudaStreamCaptureStatusNone) lanes = 1;
        (void)cudaGetLastError();
    }
    if (lanes == 2) {
        lane_stream[1] = aux_stream(1);
        if (!lane_stream[1] || cudaEventCreateWithFlags(&fork, cudaEventDisableTiming) != cudaSuccess ||
            cudaEventCreateWithFlags(&join, cudaEventDisableTiming) != cudaSuccess) {
            if (fork) cudaEventDestroy(fork);
            fork = nullptr;
            lanes = 1;
            (void)cudaGetLastError();
        }
    }
    if (lanes == 2) {
        ADRT_CUDA_CHECK(cudaEventRecord(fork, s));
        ADRT_CUDA_CHECK(cudaStreamWaitEvent(lane_stream[1], fork, 0));
    }
    for (int l = 0; l < lanes; ++l) set_persist_window(lane_stream[l], ws, lane_elems * lanes * sizeof(T));
    const long long img_elems = (long long)n * n, sino_plane = (long long)D * n;
    int rc = ADRT_B200_OK;
    int64_t wave = 0;
    for (int64_t p0 = 0; p0 < total && rc == ADRT_B200_OK; p0 += wc.planes, ++wave) {
        const int lane = (int)(wave % lanes);
        T *lane_ws = ws + lane_elems * lane;
        T *slot[2] = {lane_ws, lane_ws + slot0};
        const int np = (int)((total - p0) < wc.planes ? (total - p0) : wc.planes);
        for (int i = 0; i < pl.npass && rc == ADRT_B200_OK; ++i) {
            const plan::Pass &p = pl.pass[i];
            PassArgs a;
            a.n = n; a.D = D; a.e = 1 << p.s; a.loge = p.s; a.next_g = p.next_g; a.d_need = p.d_need;
            a.skip_zero = p.skip_zero; a.sup_loge = p.sup_loge; a.sup_gmask = p.sup_gmask;
            a.in_pitch = p.in_pitch; a.out_pitch = p.out_pitch;
            a.planes = np;
            a.x_off = 0;
            a.y_off = 0;
            a.q_first = q_first; a.q_count = q_count;
            a.plane0 = 0;
            a.side_idx = lane == 0 ? 0 : 2;
            const T *src;
            T *dst;
            if (p.src_buf < 0) {
                // caller's input: images, a public-layout sinogram, or R-layout rows (plan built with rows_in)
                const long long in_plane = p.in_pitch ? (long long)n * p.in_pitch : sino_plane;
                if (kForward) { src = in; a.plane0 = (int)p0; a.src_plane_stride = img_elems; }
                else { src = in + p0 * in_plane; a.src_plane_stride = in_plane; a.sub_delta = sub_delta; }
            } else {
                src = slot[p.src_buf];
                a.src_plane_stride = (long long)n * p.in_pitch;
            }
            if (p.dst_buf < 0) {
                // caller's output: public layout, or R-layout rows (plan built with rows_out)
                const long long out_plane = p.out_pitch ? (long long)n * p.out_pitch : sino_plane;
                dst = out + p0 * out_plane;
                a.dst_plane_stride = out_plane;
            } else {
                dst = slot[p.dst_buf];
                a.dst_plane_stride = (long long)n * p.out_pitch;
            }
            rc = dispatch_pass<T, kForward>(p, src, dst, a, lane_stream[lane]);
        }
    }
    if (lanes == 2) {
        cudaEventRecord(join, lane_stream[1]);
        cudaStreamWaitEvent(s, join, 0);
        cudaEventDestroy(fork);
        cudaEventDestroy(join);
    }
    return rc;
}

// ---- angle-block sharding (fused_plan.h part_*) ----------------------------------------------
// phase 0: the passes before the exchange (forward: all but the last, on the rank's blocks; transposed:
// the first, on the rank's columns) writing the exchange buffer `xbuf` (planes x n rows x pitch);
// phase 1: the passes after it reading `xbuf`.
template <typename T, bool kForward>
int run_part(const plan::Plan &pl, const T *in, T *out, T *xbuf, int64_t total, int q_first, int q_count, int part, int parts,
             int phase, T *ws, size_t ws_elems, cudaStream_t s)
{
    const int n = pl.n, D = pl.D, np = pl.npass;
    const int xi = kForward ? np - 2 : 0;
    const size_t slot0 = pl.ws_slot_elems[0] * (size_t)total, slot1 = pl.ws_slot_elems[1] * (size_t)total;
    if (slot0 + slot1 > ws_elems) {
        set_error("sharded transform workspace too small: need %zu elements, got %zu", slot0 + slot1, ws_elems);
        return ADRT_B200_EWORKSPACE;
    }
    T *slot[2] = {ws, ws + slot0};
    const long long img_elems = (long long)n * n, sino_plane = (long long)D * n;
    const int first = phase == 0 ? 0 : xi + 1, last = phase == 0 ? xi : np - 1;
    int rc = ADRT_B200_OK;
    for (int i = first; i <= last && rc == ADRT_B200_OK; ++i) {
        const plan::Pass &p = pl.pass[i];
        const bool angle_pass = kForward ? (i == np - 1) : (i == 0);
        const plan::YRange yr = angle_pass ? plan::part_angle_range(p, part, parts) : plan::part_block_range(p, n, part, parts);
        plan::Pass q = p;
        q.grid_y = yr.y_cnt;
        PassArgs a;
        a.n = n; a.D = D; a.e = 1 << p.s; a.loge = p.s; a.next_g = p.next_g; a.d_need = p.d_need;
            a.skip_zero = p.skip_zero; a.sup_loge = p.sup_loge; a.sup_gmask = p.sup_gmask;
        a.in_pitch = p.in_pitch; a.out_pitch = p.out_pitch;
        a.planes = (int)total;
        a.x_off = 0;
        a.y_off = yr.y_off;
        a.q_first = q_first; a.q_count = q_count;
        a.plane0 = 0;
        a.side_idx = 0;
        const T *src;
        T *dst;
        if (p.src_buf < 0) { src = in; a.src_plane_stride = kForward ? img_elems : sino_plane; }
        else { src = (i == xi + 1) ? xbuf : slot[p.src_buf]; a.src_plane_stride = (long long)n * p.in_pitch; }
        if (p.dst_buf < 0) { dst = out; a.dst_plane_stride = sino_plane; }
        else { dst = (i == xi) ? xbuf : slot[p.dst_buf]; a.dst_plane_stride = (long long)n * p.out_pitch; }
        rc = dispatch_pass<T, kForward>(q, src, dst, a, s);
    }
    return rc;
}

template <typename T, bool kForward>
size_t plan_workspace_elems(const plan::Plan &pl, int64_t B, int q_count)
{
    const size_t plane_ws = pl.ws_slot_elems[0] + pl.ws_slot_elems[1];
    const WaveCfg wc = wave_config(B * q_count, q_count, plane_ws * sizeof(T), overlap_default<kForward>(pl));
    size_t need = plane_ws * (size_t)wc.planes * wc.lanes;
    // room for the co-scheduled variant too (decided per call; one slot + counters), so that a
    // workspace sized by the query serves either
    if (pl.npass == 2) {
        const size_t total = (size_t)(B * q_count);
        const size_t co = ((pl.ws_slot_elems[0] * total + 63) & ~size_t(63)) +
                          (cosched_counter_words((int64_t)total) * sizeof(unsigned) + sizeof(T) - 1) / sizeof(T);
        if (co > need) need = co;
    }
    return need;
}

}  // namespace

template <typename T>
size_t fused_adrt_workspace_elems(int64_t B, int64_t n, int q_count)
{
    plan::Plan pl;
    if (n > kMaxN || !plan::make_forward_plan(n, sizeof(T), &pl)) return (size_t)-1;
    return plan_workspace_elems<T, true>(pl, B, q_count);
}

template <typename T>
size_t fused_bdrt_workspace_elems(int64_t B, int64_t n, int q_count)
{
    plan::Plan pl;
    if (n > kMaxN || !plan::make_transposed_plan(n, sizeof(T), &pl)) return (size_t)-1;
    return plan_workspace_elems<T, false>(pl, B, q_count);
}

template <typename T>
int fused_adrt(const T *in, T *out, int64_t B, int64_t n, int q_first, int q_count, T *ws, size_t ws_elems,
               cudaStream_t s, bool *handled, bool rows_out)
{
    // rows_out: `out` receives R-layout rows, planes x n x round4(2n-1) elements (fused_plan.h)
    plan::Plan pl;
    *handled = false;
    if (n > kMaxN || !plan::make_forward_plan(n, sizeof(T), &pl, rows_out)) return ADRT_B200_OK;
    *handled = true;
    return run_plan<T, true>(pl, in, out, B, q_first, q_count, ws, ws_elems, s);
}

// Can the first pass of the transposed plan subtract a second sinogram while it loads (fused_bdrt's `sub`)?
// Only the fused_tile.h kernels whose loader runs the first radix-4 step on the public layout do.
inline bool sub_on_load_ok(const plan::Plan &pl)
{
    const plan::Pass &p = pl.pass[0];
    return pl.npass >= 2 && !p.stream && !p.staged && p.src_buf < 0 && p.in_pitch == 0 && p.load == tile::LOAD_QCOLS &&
           p.store == tile::STORE_WROWS && p.M >= 4 && p.M <= 6;
}

template <typename T>
bool fused_bdrt_sub_ok(int64_t n, int64_t rows)
{
    plan::Plan pl;
    if (n > kMaxN || !plan::make_transposed_plan(n, sizeof(T), &pl, rows, false)) return false;
    return sub_on_load_ok(pl);
}

template <typename T>
int fused_bdrt(const T *in, T *out, int64_t B, int64_t n, int q_count, int64_t rows, T *ws, size_t ws_elems,
               cudaStream_t s, bool *handled, bool rows_in, const T *sub)
{
    // rows < 2n-1: only offsets d < rows of every output plane are computed, the rest of
    // `out` is left as it was
    plan::Plan pl;
    *handled = false;
    // rows_in: `in` holds R-layout rows as written by fused_adrt(..., rows_out = true)
    if (n > kMaxN || !plan::make_transposed_plan(n, sizeof(T), &pl, rows, rows_in)) return ADRT_B200_OK;
    *handled = true;
    long long sub_delta = 0;
    if (sub) {
        if (rows_in || !sub_on_load_ok(pl)) {
            set_error("internal: this transposed plan has no subtract-on-load first pass (n = %lld)", (long long)n);
            return ADRT_B200_EINVAL;
        }
        sub_delta = (long long)(reinterpret_cast<const char *>(sub) - reinterpret_cast<const char *>(in));
    }
    return run_plan<T, false>(pl, in, out, B, 0, q_count, ws, ws_elems, s, sub_delta);
}

template <typename T>
bool make_part_plan(int64_t n, int parts, int m_last, bool forward, int64_t rows, plan::Plan *pl)
{
    if (n > kMaxN || parts < 2 || (parts & (parts - 1)) || m_last < 1 || (1 << m_last) < parts) return false;
    const int K = plan::ilog2(n);
    const std::vector<int> ms = plan::part_split(K, sizeof(T), m_last);
    if (ms.size() < 2) return false;
    return forward ? plan::make_forward_plan_split(n, sizeof(T), ms, pl) : plan::make_transposed_plan_split(n, sizeof(T), ms, pl, rows);
}

template <typename T>
size_t part_exchange_pitch(int64_t n, int m_last, bool forward)
{
    plan::Plan pl;
    if (!make_part_plan<T>(n, 2, m_last, forward, -1, &pl)) return 0;
    return (size_t)pl.pass[forward ? pl.npass - 2 : 0].out_pitch;
}

// leading elements of every bdrt exchange row that the passes after the exchange read when only offsets
// d < rows of the result are wanted (the rest of the row need not travel)
template <typename T>
size_t part_exchange_cols(int64_t n, int m_last, int64_t rows)
{
    plan::Plan pl;
    if (!make_part_plan<T>(n, 2, m_last, false, rows, &pl)) return 0;
    const long long pitch = pl.pass[0].out_pitch, need = plan::round4(pl.pass[0].d_need);
    return (size_t)(need < pitch ? need : pitch);
}

template <typename T>
size_t part_workspace_elems(int64_t planes, int64_t n, int m_last)
{
    plan::Plan f, b;
    if (!make_part_plan<T>(n, 2, m_last, true, -1, &f) || !make_part_plan<T>(n, 2, m_last, false, -1, &b)) return (size_t)-1;
    const size_t wf = (f.ws_slot_elems[0] + f.ws_slot_elems[1]) * (size_t)planes;
    const size_t wb = (b.ws_slot_elems[0] + b.ws_slot_elems[1]) * (size_t)planes;
    return wf > wb ? wf : wb;
}

template <typename T>
int fused_adrt_part(const T *img, T *xbuf, T *sino, int64_t B, int64_t n, int q_first, int q_count, int part, int parts,
                    int m_last, int phase, T *ws, size_t ws_elems, cudaStream_t s)
{
    plan::Plan pl;
    if (!make_part_plan<T>(n, parts, m_last, true, -1, &pl)) {
        set_error("no sharded forward plan for n=%lld parts=%d m_last=%d", (long long)n, parts, m_last);
        return ADRT_B200_EINVAL;
    }
    return run_part<T, true>(pl, img, sino, xbuf, B * q_count, q_first, q_count, part, parts, phase, ws, ws_elems, s);
}

template <typename T>
int fused_bdrt_part(const T *sino, T *xbuf, T *out, int64_t planes, int64_t n, int64_t rows, int part, int parts, int m_last,
                    int phase, T *ws, size_t ws_elems, cudaStream_t s)
{
    plan::Plan pl;
    if (!make_part_plan<T>(n, parts, m_last, false, rows, &pl)) {
        set_error("no sharded transposed plan for n=%lld parts=%d m_last=%d", (long long)n, parts, m_last);
        return ADRT_B200_EINVAL;
    }
    return run_part<T, false>(pl, sino, out, xbuf, planes, 0, 1, part, parts, phase, ws, ws_elems, s);
}

#define INSTANTIATE(T)                                                      \
    template size_t part_exchange_pitch<T>(int64_t, int, bool);             \
    template size_t part_exchange_cols<T>(int64_t, int, int64_t);           \
    template size_t part_workspace_elems<T>(int64_t, int64_t, int);         \
    template int fused_adrt_part<T>(const T *, T *, T *, int64_t, int64_t, int, int, int, int, int, int, T *, size_t, cudaStream_t); \
    template int fused_bdrt_part<T>(const T *, T *, T *, int64_t, int64_t, int64_t, int, int, int, int, T *, size_t, cudaStream_t); \
    template size_t fused_adrt_workspace_elems<T>(int64_t, int64_t, int);   \
    template size_t fused_bdrt_workspace_elems<T>(int64_t, int64_t, int);   \
    template int fused_adrt<T>(const T *, T *, int64_t, int64_t, int, int, T *, size_t, cudaStream_t, bool *, bool); \
    template int fused_bdrt<T>(const T *, T *, int64_t, int64_t, int, int64_t, T *, size_t, cudaStream_t, bool *, bool, const T *); \
    template bool fused_bdrt_sub_ok<T>(int64_t, int64_t);
INSTANTIATE(float)
INSTANTIATE(double)

}  // namespace adrt_b200
