// Staged five-stage fp32 passes for sm_100a (stage_tile.h): persistent CTAs, input tiles brought in
// by tensor-map / bulk copies into a staging buffer, the copies of a CTA's next tile in flight while
// the second butterfly step and the store of the current one run.
#include "pass_args.h"
#include "stage_tile.h"

#include <cuda.h>

#include <mutex>

namespace adrt_b200 {

namespace {

using sgtile::TmaMap;

// The CTA's tiles: items blockIdx.x, blockIdx.x + gridDim.x, ... of the launch's
// planes x groups x d-tiles (d-tile fastest: CTAs that run together read neighbouring tiles).
struct Item {
    unsigned it;
    int plane;      // plane of the launch (transposed) / image index (forward)
    int oplane;     // plane of the launch: which destination plane the tile stores to
    int mode;
    tile::TileCtx c;
};

template <typename Prog>
__device__ __forceinline__ bool find_item(unsigned it, unsigned total, unsigned nx, unsigned ny, const PassArgs &a, Item &o)
{
    for (; it < total; it += gridDim.x) {
        const unsigned x = it % nx, r = it / nx, y = r % ny, plane = r / ny;
        o.c.g = (int)y + a.y_off;
        o.c.k0 = o.c.g >> a.loge;
        o.c.a_g = o.c.g & (a.e - 1);
        o.c.d0 = ((int)x + a.x_off) * Prog::TD;
        const int mode = Prog::classify(o.c);
        if (!Prog::runs(mode) || (mode == tile::TILE_ZERO && a.skip_zero)) continue;
        o.it = it;
        o.mode = mode;
        o.oplane = (int)plane;
        if (Prog::kImage) {
            const int gp = (int)plane + a.plane0;
            o.c.q = a.q_first + gp % a.q_count;
            o.plane = gp / a.q_count;
        } else {
            o.c.q = 0;
            o.plane = (int)plane;
        }
        return true;
    }
    o.it = total;
    return false;
}

template <typename Prog>
__global__ void __launch_bounds__(Prog::NT, Prog::MIN_CTAS)
staged_kernel(const __grid_constant__ TmaMap tm, const float *__restrict__ src, float *__restrict__ dst, PassArgs a,
              unsigned nx, unsigned ny, unsigned total)
{
    extern __shared__ __align__(128) unsigned char smem_raw[];
    __shared__ unsigned long long bulk_bar;
    // tensor copies want a 128-byte aligned destination
    const unsigned mis = (unsigned)__cvta_generic_to_shared(smem_raw) & 127u;
    float *in = reinterpret_cast<float *>(smem_raw + ((128u - mis) & 127u));   // receives the tile's input, ends up holding its output
    float *mid = in + sgtile::STG_FLOATS;                                      // between the two butterfly steps
    const int tid = threadIdx.x;

    Item cur, nxt;
    cur.c.n = a.n;
    cur.c.D = a.D;
    cur.c.e = a.e;
    cur.c.next_g = a.next_g;
    cur.c.d_need = a.d_need;
    cur.c.sup_loge = a.sup_loge;
    cur.c.sup_gmask = a.sup_gmask;
    cur.c.in_pitch = a.in_pitch;
    cur.c.out_pitch = a.out_pitch;
    cur.c.q = 0;
    nxt = cur;

    typename Prog::State st;
    stile::bulk_init(st.bar, &bulk_bar, Prog::NT, tid);

    if (!find_item<Prog>(blockIdx.x, total, nx, ny, a, cur)) return;
    unsigned loading = total;   // item whose copies are in flight towards (or have landed in) `in`
    while (cur.it < total) {
        find_item<Prog>(cur.it + gridDim.x, total, nx, ny, a, nxt);
        float *dp = dst + (long long)cur.oplane * a.dst_plane_stride;
        if (cur.mode == tile::TILE_ZERO) {
            Prog::zero_tile(mid, dp, cur.c, tid);
        } else {
            if (loading != cur.it) Prog::template phase_ct<0>(cur.mode, in, mid, st, tm, src, dp, cur.c, cur.plane, tid);
            Prog::template phase_ct<1>(cur.mode, in, mid, st, tm, src, dp, cur.c, cur.plane, tid);
            __syncthreads();
            Prog::template phase_ct<2>(cur.mode, in, mid, st, tm, src, dp, cur.c, cur.plane, tid);
            __syncthreads();
            Prog::template phase_ct<3>(cur.mode, in, mid, st, tm, src, dp, cur.c, cur.plane, tid);
            __syncthreads();
            Prog::template phase_ct<4>(cur.mode, in, mid, st, tm, src, dp, cur.c, cur.plane, tid);
            // `mid` is free, `in` drains through the bulk stores: the buffers swap roles, and the next
            // tile's copies run while the stores drain (and under the other CTAs of the SM)
            float *t = in;
            in = mid;
            mid = t;
            if (nxt.it < total && nxt.mode != tile::TILE_ZERO) {
                Prog::template phase_ct<0>(nxt.mode, in, mid, st, tm, src, dp, nxt.c, nxt.plane, tid);
                loading = nxt.it;
            }
        }
        cur = nxt;
    }
    sgtile::bulk_store_wait_read();   // shared memory must outlive the reads of the last bulk stores
}

// ---- tensor maps -------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                                  const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn encode_fn()
{
    static std::mutex mu;
    static EncodeTiledFn fn = nullptr;
    std::lock_guard<std::mutex> lock(mu);
    if (!fn) {
        void *p = nullptr;
        cudaDriverEntryPointQueryResult q = cudaDriverEntryPointSymbolNotFound;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
            q == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<EncodeTiledFn>(p);
        else
            (void)cudaGetLastError();
    }
    return fn;
}

// rank-3 fp32 tensor: `cols` contiguous columns, `rows` rows `row_stride` elements apart, `planes`
// planes `plane_stride` elements apart; box = SW x BOX_ROWS x 1, out-of-range elements read as 0
int make_map(TmaMap *m, const float *base, long long cols, long long rows, long long planes, long long row_stride,
             long long plane_stride)
{
    static_assert(sizeof(CUtensorMap) <= sizeof(m->opaque), "tensor map does not fit");
    EncodeTiledFn enc = encode_fn();
    if (!enc) {
        set_error("cuTensorMapEncodeTiled is not available from this driver");
        return ADRT_B200_ECUDA;
    }
    m->base = base;
    m->dim0 = (int)cols; m->dim1 = (int)rows; m->dim2 = (int)planes;
    m->stride1 = row_stride; m->stride2 = plane_stride;
    const cuuint64_t gdim[3] = {(cuuint64_t)cols, (cuuint64_t)rows, (cuuint64_t)planes};
    const cuuint64_t gstr[2] = {(cuuint64_t)row_stride * 4, (cuuint64_t)plane_stride * 4};
    const cuuint32_t box[3] = {(cuuint32_t)sgtile::SW, (cuuint32_t)sgtile::BOX_ROWS, 1};
    const cuuint32_t est[3] = {1, 1, 1};
    const CUresult r = enc(reinterpret_cast<CUtensorMap *>(m->opaque), CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3,
                           const_cast<float *>(base), gdim, gstr, box, est, CU_TENSOR_MAP_INTERLEAVE_NONE,
                           CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        set_error("cuTensorMapEncodeTiled failed (%d) for a %lld x %lld x %lld tensor", (int)r, cols, rows, planes);
        return ADRT_B200_ECUDA;
    }
    return ADRT_B200_OK;
}

int resident_ctas(int per_sm)
{
    // ADRT_B200_STAGE_CTAS: persistent grid size (tuning)
    if (const char *e = getenv("ADRT_B200_STAGE_CTAS"))
        if (atoi(e) > 0) return atoi(e);
    int dev = 0, sms = 148;
    if (cudaGetDevice(&dev) != cudaSuccess || cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess) {
        (void)cudaGetLastError();
        sms = 148;
    }
    return sms * per_sm;
}

template <typename Prog>
int launch(const TmaMap &tm, const float *src, float *dst, PassArgs a, int x_first, int tiles_x, int tiles_y, int max_ctas,
           cudaStream_t s)
{
    if (tiles_x <= 0 || tiles_y <= 0 || a.planes <= 0) return ADRT_B200_OK;
    a.x_off = x_first;
    const double items = (double)a.planes * tiles_x * tiles_y;
    if (items >= 4.0e9) {
        set_error("staged pass: %g tiles exceed the 32-bit work index", items);
        return ADRT_B200_EINVAL;
    }
    const unsigned total = (unsigned)a.planes * (unsigned)tiles_x * (unsigned)tiles_y;
    auto kern = staged_kernel<Prog>;
    // the two tile buffers (+ 32 bytes: the top segment of a transposed step reads a few cells past the
    // last row) + room to align them
    const size_t smem = (size_t)(2 * sgtile::STG_FLOATS) * sizeof(float) + 32 + 128;
    ADRT_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    unsigned ctas = (unsigned)max_ctas;
    if (ctas > total) ctas = total;
    kern<<<ctas, Prog::NT, smem, s>>>(tm, src, dst, a, (unsigned)tiles_x, (unsigned)tiles_y, total);
    ADRT_LAUNCH_CHECK();
    return ADRT_B200_OK;
}

// interior tiles and the tiles that reach offset D as two launches (the latter beside the former on a
// helper stream), as launch_bwd in stream_adrt.cu
int launch_bwd5(const plan::Pass &p, const float *src, float *dst, const PassArgs &a, cudaStream_t s)
{
    typedef sgtile::BwdStaged<5, false> Inner;
    typedef sgtile::BwdStaged<5, true> Edge;
    TmaMap tm;
    int rc = make_map(&tm, src, a.n, a.D, a.planes, a.n, a.src_plane_stride);
    if (rc != ADRT_B200_OK) return rc;
    int xm = Inner::first_masked_tile(a.D);
    if (xm > p.grid_x) xm = p.grid_x;
    const int nmask = p.grid_x - xm;
    const int ctas = resident_ctas(Inner::MIN_CTAS);
    cudaStream_t side = (xm > 0 && nmask > 0) ? aux_stream(a.side_idx) : nullptr;
    cudaStreamCaptureStatus cap = cudaStreamCaptureStatusNone;
    if (side && cudaStreamIsCapturing(s, &cap) == cudaSuccess && cap != cudaStreamCaptureStatusNone) side = nullptr;
    (void)cudaGetLastError();
    cudaEvent_t fork = nullptr, join = nullptr;
    if (side) {
        if (cudaEventCreateWithFlags(&fork, cudaEventDisableTiming) != cudaSuccess ||
            cudaEventCreateWithFlags(&join, cudaEventDisableTiming) != cudaSuccess) {
            if (fork) cudaEventDestroy(fork);
            side = nullptr;
            (void)cudaGetLastError();
        }
    }
    if (side) {
        ADRT_CUDA_CHECK(cudaEventRecord(fork, s));
        ADRT_CUDA_CHECK(cudaStreamWaitEvent(side, fork, 0));
        rc = launch<Edge>(tm, src, dst, a, xm, nmask, p.grid_y, ctas / 4 > 0 ? ctas / 4 : 1, side);
        if (rc == ADRT_B200_OK) rc = launch<Inner>(tm, src, dst, a, 0, xm, p.grid_y, ctas, s);
        cudaEventRecord(join, side);
        cudaStreamWaitEvent(s, join, 0);
        cudaEventDestroy(fork);
        cudaEventDestroy(join);
        return rc;
    }
    rc = launch<Inner>(tm, src, dst, a, 0, xm, p.grid_y, ctas, s);
    if (rc != ADRT_B200_OK) return rc;
    return launch<Edge>(tm, src, dst, a, xm, nmask, p.grid_y, ctas, s);
}

int launch_fwd5(const plan::Pass &p, const float *src, float *dst, const PassArgs &a, cudaStream_t s)
{
    typedef sgtile::FwdStaged<5> Prog;
    // images of the whole batch behind `src`; the launch's planes start at plane0
    const long long images = ((long long)a.plane0 + a.planes + a.q_count - 1) / a.q_count;
    TmaMap tm;
    const int rc = make_map(&tm, src, a.n, a.n, images, a.n, a.src_plane_stride);
    if (rc != ADRT_B200_OK) return rc;
    return launch<Prog>(tm, src, dst, a, 0, p.grid_x, p.grid_y, resident_ctas(Prog::MIN_CTAS), s);
}

}  // namespace

bool staged_pass_available() { return encode_fn() != nullptr; }

int launch_staged_pass(const plan::Pass &p, bool forward, const float *src, float *dst, const PassArgs &a, cudaStream_t s)
{
    if (p.M == 5 && forward && p.load == tile::LOAD_IMAGE && p.store == tile::STORE_WROWS) return launch_fwd5(p, src, dst, a, s);
    if (p.M == 5 && !forward && p.load == tile::LOAD_QCOLS && p.store == tile::STORE_WROWS) return launch_bwd5(p, src, dst, a, s);
    set_error("internal: no staged kernel for M=%d forward=%d kinds %d/%d", p.M, (int)forward, p.load, p.store);
    return ADRT_B200_EINVAL;
}

}  // namespace adrt_b200
