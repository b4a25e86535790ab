/* adrt_b200 -- C ABI of the B200-native ADRT engine (sm_100a).
 *
 * This header is the drop-in boundary: one entry point per function of the
 * reference's native module `adrt._adrt_cdefs`
 * (/root/reference/src/adrt/adrt_cdefs_py.cpp:852-863).  Everything is plain
 * C: raw pointers, int64 sizes, a dtype enum, an opaque cudaStream_t.  No
 * Python, no torch types, no exceptions.  Every function returns 0 on
 * success and a non-zero adrt_b200_status otherwise; adrt_b200_last_error()
 * gives the message for the calling thread.
 *
 * Array layouts are the reference's public ones (C order):
 *   image      (B, n, n)
 *   sinogram   (B, 4, 2n-1, n)         "ADRT output shape"
 *   cartesian  (B, n, 4n)
 * n must be a power of two (adrt_cdefs_common.cpp:176-192).  Shape validity
 * is re-checked here (status ADRT_B200_EINVAL) but the Python-visible error
 * messages are produced by the shim above this ABI (adrt_b200/_adrt_cdefs.py).
 *
 * Two families:
 *   adrt_b200_<op>(...)       device pointers, asynchronous on `stream`,
 *                             caller-provided workspace (size from
 *                             adrt_b200_<op>_workspace_bytes), no allocation,
 *                             no synchronisation.  Device pointers (input,
 *                             output, workspace) must be 32-byte aligned --
 *                             the kernels use 16/32-byte vector accesses;
 *                             cudaMalloc and framework allocators satisfy
 *                             this, a misaligned pointer is ADRT_B200_EINVAL.
 *   adrt_b200_host_<op>(...)  host pointers (pageable or pinned); performs the
 *                             H2D / compute / D2H pipeline in batch chunks on
 *                             device `device` and returns when the output is
 *                             complete.  This is what the NumPy path calls.
 */
#ifndef ADRT_B200_H
#define ADRT_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef enum {
    ADRT_B200_F32 = 0,
    ADRT_B200_F64 = 1
} adrt_b200_dtype;

typedef enum {
    ADRT_B200_OK = 0,
    ADRT_B200_EINVAL = 1,    /* bad shape / dtype / step / null pointer          */
    ADRT_B200_EWORKSPACE = 2,/* workspace too small                              */
    ADRT_B200_ECUDA = 3,     /* a CUDA runtime call failed (see last_error)      */
    ADRT_B200_ENOMEM = 4     /* host or device allocation failed                 */
} adrt_b200_status;

/* Library / device info ---------------------------------------------------- */
int         adrt_b200_version(void);              /* 10000*major+100*minor+patch */
const char *adrt_b200_last_error(void);           /* thread-local, never NULL    */
int         adrt_b200_device_count(void);         /* <0 on CUDA error            */
/* Tuning knob for tests / benchmarks: 0 = automatic (fused multi-stage
 * kernels), 1 = force the one-kernel-per-stage path for adrt/bdrt.           */
void        adrt_b200_set_mode(int mode);
int         adrt_b200_get_mode(void);
/* Number of kernels this library has launched since load (all threads).     */
int64_t     adrt_b200_launch_count(void);

/* Pure shape helpers (adrt_cdefs_common.cpp:140-142, 176-330) -------------- */
int         adrt_b200_num_iters(int64_t n);

/* Device-pointer family ---------------------------------------------------- */

/* adrt.adrt: adrt_cdefs_py.cpp:275-341 -> adrt_basic (adrt_cdefs_adrt.hpp:101-212).
 * in (B,n,n) -> out (B,4,2n-1,n). */
size_t adrt_b200_adrt_workspace_bytes(int64_t B, int64_t n, int dtype);
int    adrt_b200_adrt(const void *in, void *out, int64_t B, int64_t n, int dtype,
                      void *workspace, size_t workspace_bytes, void *stream);

/* adrt.bdrt: adrt_cdefs_py.cpp:482-548 -> bdrt_basic (adrt_cdefs_bdrt.hpp:121-187).
 * in (B,4,2n-1,n) -> out same shape. */
size_t adrt_b200_bdrt_workspace_bytes(int64_t B, int64_t n, int dtype);
int    adrt_b200_bdrt(const void *in, void *out, int64_t B, int64_t n, int dtype,
                      void *workspace, size_t workspace_bytes, void *stream);

/* Plane subsets, for sharding ONE large image over several GPUs: the
 * four quadrants of adrt are independent transforms of four orientations of the
 * image (adrt_cdefs_adrt.hpp:124-175), and bdrt treats every (2n-1, n) plane
 * independently.  adrt_quadrants: in (B,n,n) -> out (B,q_count,2n-1,n) holding
 * quadrants q_first .. q_first+q_count-1;  bdrt_planes: in/out (planes,2n-1,n). */
size_t adrt_b200_adrt_quadrants_workspace_bytes(int64_t B, int64_t n, int dtype, int q_count);
int    adrt_b200_adrt_quadrants(const void *in, void *out, int64_t B, int64_t n, int dtype,
                                int q_first, int q_count, void *workspace, size_t workspace_bytes,
                                void *stream);
size_t adrt_b200_bdrt_planes_workspace_bytes(int64_t planes, int64_t n, int dtype);
int    adrt_b200_bdrt_planes(const void *in, void *out, int64_t planes, int64_t n, int dtype,
                             void *workspace, size_t workspace_bytes, void *stream);
/* bdrt whose caller keeps only offsets d < rows of every plane -- utils.truncate keeps
 * d < n (utils.py:231-242), which is all that the normal operator truncate(bdrt(adrt(x)))
 * of the CG recipe (docs/examples.cginverse.md:45-52) and iadrt_fmg_step (core.py:318-331)
 * ever read.  Rows d < rows of `out` are bit-identical to adrt_b200_bdrt_planes; the other
 * rows of `out` are unspecified.  Workspace: adrt_b200_bdrt_planes_workspace_bytes. */
int    adrt_b200_bdrt_rows(const void *in, void *out, int64_t planes, int64_t n, int64_t rows,
                           int dtype, void *workspace, size_t workspace_bytes, void *stream);

/* Fused normal operator A^T A of the CG recipe (docs/examples.cginverse.md:45-52
 * ADRTNormalOperator._matmat: mean over quadrants of truncate(bdrt(adrt(x))); the same expression
 * with a divisor is the gradient term of iadrt_fmg_step, core.py:329):
 *   out (B,n,n) = mean_q(truncate(bdrt(adrt(in (B,n,n)))) / divisor)
 * The sinogram is handed from adrt to bdrt as workspace rows (never stored in, nor re-loaded from, the
 * public (d, column) layout) and the back-projection stops at offset n; bit-identical to composing
 * adrt_b200_adrt, adrt_b200_bdrt and adrt_b200_truncate_mean.
 * adrt_bdrt_rows: the same pipeline without the quadrant mean, for quadrants q_first ..
 * q_first+q_count-1: out (B,q_count,2n-1,n) whose offsets d < n equal bdrt(adrt(in)); n >= 2. */
size_t adrt_b200_normal_operator_workspace_bytes(int64_t B, int64_t n, int dtype);
int    adrt_b200_normal_operator(const void *in, void *out, int64_t B, int64_t n, double divisor,
                                 int dtype, void *workspace, size_t workspace_bytes, void *stream);
size_t adrt_b200_adrt_bdrt_rows_workspace_bytes(int64_t B, int64_t n, int dtype, int q_count);
int    adrt_b200_adrt_bdrt_rows(const void *in, void *out, int64_t B, int64_t n, int dtype,
                                int q_first, int q_count, void *workspace, size_t workspace_bytes,
                                void *stream);

/* Angle-block sharding of ONE large image over `parts` (2, 4, 8) ranks per quadrant (SURVEY.md 8e row 2;
 * the structure is that of adrt_core, adrt_cdefs_adrt.hpp:55-96: a stage only pairs rows of adjacent
 * blocks that have the SAME incoming angle).  The last m_last stages (2^m_last >= parts) are one fused
 * pass.  adrt_part phase 0 runs every earlier pass on the rank's own image-row blocks and writes the
 * exchange buffer xbuf = (planes, 2^m_last blocks, n / 2^m_last angles, pitch) with the rows (blk, a) of
 * its blocks; the ranks then exchange rows (the owner of block blk sends row (blk, a) to the owner of
 * angle a: adrt_b200/_shard.py, one batched NCCL send/recv); phase 1 runs the last pass for the rank's
 * angles and fills columns [part*n/parts, (part+1)*n/parts) of every (2n-1, n) plane of `sino`.
 * bdrt_part is the mirror image: phase 0 reads those columns of `sino` and writes rows (blk, a) of the
 * rank's angles, the exchange goes the other way, phase 1 fills the rank's columns of `out` (offsets
 * d < rows only; rows = -1: all).  Results are bit-identical to adrt_b200_adrt / _bdrt.
 * part_exchange_pitch: elements per xbuf row (forward != 0: adrt, else bdrt); 0 = unsupported shape. */
size_t adrt_b200_part_exchange_pitch(int64_t n, int dtype, int m_last, int forward);
/* leading elements of every bdrt xbuf row that phase 1 reads when only offsets d < rows are wanted
 * (rows = -1: all): the exchange need not move the rest of the row */
size_t adrt_b200_part_exchange_cols(int64_t n, int dtype, int m_last, int64_t rows);
size_t adrt_b200_part_workspace_bytes(int64_t planes, int64_t n, int dtype, int m_last);
int    adrt_b200_adrt_part(const void *img, void *xbuf, void *sino, int64_t B, int64_t n, int dtype,
                           int q_first, int q_count, int part, int parts, int m_last, int phase,
                           void *workspace, size_t workspace_bytes, void *stream);
int    adrt_b200_bdrt_part(const void *sino, void *xbuf, void *out, int64_t planes, int64_t n,
                           int64_t rows, int dtype, int part, int parts, int m_last, int phase,
                           void *workspace, size_t workspace_bytes, void *stream);

/* adrt.core.adrt_step / bdrt_step: adrt_cdefs_py.cpp:343-412, 550-619 ->
 * adrt_step (adrt_cdefs_adrt.hpp:215-258), bdrt_step (adrt_cdefs_bdrt.hpp:190-244).
 * 0 <= step < num_iters(n).  in/out (B,4,2n-1,n), must not alias. */
int    adrt_b200_adrt_step(const void *in, void *out, int64_t B, int64_t n, int step,
                           int dtype, void *stream);
int    adrt_b200_bdrt_step(const void *in, void *out, int64_t B, int64_t n, int step,
                           int dtype, void *stream);

/* adrt.core.adrt_init (core.py:123-176; pure NumPy in the reference):
 * in (B,n,n) -> out (B,4,2n-1,n). */
int    adrt_b200_adrt_init(const void *in, void *out, int64_t B, int64_t n, int dtype,
                           void *stream);

/* adrt.iadrt: adrt_cdefs_py.cpp:414-480 -> iadrt_basic (adrt_cdefs_iadrt.hpp:110-176). */
size_t adrt_b200_iadrt_workspace_bytes(int64_t B, int64_t n, int dtype);
int    adrt_b200_iadrt(const void *in, void *out, int64_t B, int64_t n, int dtype,
                       void *workspace, size_t workspace_bytes, void *stream);

/* Press FMG operators: adrt_cdefs_py.cpp:684-850 -> adrt_cdefs_fmg.hpp:53-171.
 * restriction (B,4,2n-1,n) -> (B,4,n-1,n/2), n even >= 2;
 * prolongation (B,h,w) -> (B,2h,2w); highpass (B,h,w) -> same, h,w >= 2. */
int    adrt_b200_fmg_restriction(const void *in, void *out, int64_t B, int64_t n, int dtype,
                                 void *stream);
int    adrt_b200_fmg_prolongation(const void *in, void *out, int64_t B, int64_t h, int64_t w,
                                  int dtype, void *stream);
int    adrt_b200_fmg_highpass(const void *in, void *out, int64_t B, int64_t h, int64_t w,
                              int dtype, void *stream);

/* adrt.core.iadrt_fmg_step (core.py:265-331; pure Python over the operators above in
 * the reference): one full-multigrid pass, every level on the device in one call.
 * in (B,4,2n-1,n) -> out (B,n,n); bit-identical to composing restriction /
 * prolongation / adrt / bdrt / truncate / mean / highpass as core.py:318-331 does. */
size_t adrt_b200_fmg_step_workspace_bytes(int64_t B, int64_t n, int dtype);
int    adrt_b200_fmg_step(const void *in, void *out, int64_t B, int64_t n, int dtype,
                          void *workspace, size_t workspace_bytes, void *stream);

/* adrt.utils.interp_to_cart: adrt_cdefs_py.cpp:621-682 -> interp_adrtcart
 * (adrt_cdefs_interp_adrtcart.hpp:61-114).  The (quadrant, height, slope,
 * factor) table depends only on (n, dtype); it is computed on the host with
 * the reference's float32 libm expressions, cached per device, and the gather
 * runs on the GPU.  in (B,4,2n-1,n) -> out (B,n,4n), 2 <= n <= 2^22. */
int    adrt_b200_interp_to_cart(const void *in, void *out, int64_t B, int64_t n, int dtype,
                                void *stream);

/* Device-side glue used by the multigrid / CG drivers (the reference does
 * these in NumPy on the host: core.py:318-331, utils.py:231-242).
 *   truncate:        (B,4,2n-1,n) -> (B,4,n,n)                 utils.truncate
 *   truncate_mean:   (B,4,2n-1,n) -> (B,n,n)
 *                    = mean_q(truncate(a) / divisor), evaluated as
 *                    (((t0/div + t1/div) + t2/div) + t3/div) / 4  (NumPy order)
 *   sub / add:       out = a - b, out = a + b  (elementwise, count elements)
 *   sub_inplace:     a -= b                                                  */
int    adrt_b200_truncate(const void *in, void *out, int64_t B, int64_t n, int dtype, void *stream);
/* adrt.utils.stitch_adrt / unstitch_adrt (utils.py:111-134, 162-188; pure NumPy in the reference):
 *   stitch:   (B,4,2n-1,n) -> (B,3n-2,4n), or (B,3n-2,4n-4) with remove_repeated
 *   unstitch: the inverse, from either width (trimmed != 0: the 4n-4 form, n >= 2)            */
int    adrt_b200_stitch(const void *in, void *out, int64_t B, int64_t n, int remove_repeated,
                        int dtype, void *stream);
int    adrt_b200_unstitch(const void *in, void *out, int64_t B, int64_t n, int trimmed,
                          int dtype, void *stream);
int    adrt_b200_truncate_mean(const void *in, void *out, int64_t B, int64_t n, double divisor,
                               int dtype, void *stream);
/* truncate_mean over the all-gathered shares of a sharded back-projection (single image over several
 * GPUs, adrt_b200/_shard.py): rank r = group * parts + p contributed (B, per, n, n/parts) -- offsets
 * d < n, columns [p*n/parts, (p+1)*n/parts) of its group's `per` quadrants -- and `in` holds the
 * shares rank-major: (4 / per * parts, B, per, n, n / parts).  Same arithmetic and order as
 * truncate_mean; n / parts >= 32. */
int    adrt_b200_truncate_mean_shares(const void *in, void *out, int64_t B, int64_t n, int per, int parts,
                                      double divisor, int dtype, void *stream);
int    adrt_b200_sub(const void *a, const void *b, void *out, int64_t count, int dtype, void *stream);
int    adrt_b200_add(const void *a, const void *b, void *out, int64_t count, int dtype, void *stream);
/* Vector updates of conjugate gradients on the normal equations (the reference's recipe,
 * docs/examples.cginverse.md:40-67, hands `truncate(bdrt(adrt(x)))` to scipy.sparse.linalg.cg): one
 * iteration is the operator plus three passes over the vectors.  `state`: 4 doubles on the device,
 * [0] = r.r, [1] = p.Ap, [2] = the new r.r; `ws`: adrt_b200_cg_workspace_bytes() bytes of device scratch.
 *   cg_dot:       state[slot] = a . b
 *   cg_update:    alpha = state[0] / state[1];  x += alpha p;  r -= alpha ap;  state[2] = r . r
 *   cg_direction: beta = state[2] / state[0];   p = r + beta p;  state[0] = state[2]
 * Sums are accumulated in double in a fixed order (replicated ranks stay bit-identical). */
size_t adrt_b200_cg_workspace_bytes(void);
int    adrt_b200_cg_dot(const void *a, const void *b, void *state, int slot, int64_t count, int dtype,
                        void *ws, size_t ws_bytes, void *stream);
int    adrt_b200_cg_update(void *x, void *r, const void *p, const void *ap, void *state, int64_t count, int dtype,
                           void *ws, size_t ws_bytes, void *stream);
int    adrt_b200_cg_direction(void *p, const void *r, void *state, int64_t count, int dtype, void *stream);

/* Host-pointer family (NumPy path) ----------------------------------------- *
 * Same semantics as above with host buffers.  `device` is the CUDA ordinal.
 * Buffers may be pageable or pinned (detected with cudaPointerGetAttributes;
 * pinned buffers are copied directly, pageable ones go through an internal
 * pinned staging ring filled by several host threads).                      */
int adrt_b200_host_adrt(const void *in, void *out, int64_t B, int64_t n, int dtype, int device);
int adrt_b200_host_bdrt(const void *in, void *out, int64_t B, int64_t n, int dtype, int device);
int adrt_b200_host_iadrt(const void *in, void *out, int64_t B, int64_t n, int dtype, int device);
int adrt_b200_host_adrt_step(const void *in, void *out, int64_t B, int64_t n, int step, int dtype, int device);
int adrt_b200_host_bdrt_step(const void *in, void *out, int64_t B, int64_t n, int step, int dtype, int device);
int adrt_b200_host_adrt_init(const void *in, void *out, int64_t B, int64_t n, int dtype, int device);
int adrt_b200_host_fmg_restriction(const void *in, void *out, int64_t B, int64_t n, int dtype, int device);
int adrt_b200_host_fmg_prolongation(const void *in, void *out, int64_t B, int64_t h, int64_t w, int dtype, int device);
int adrt_b200_host_fmg_highpass(const void *in, void *out, int64_t B, int64_t h, int64_t w, int dtype, int device);
int adrt_b200_host_interp_to_cart(const void *in, void *out, int64_t B, int64_t n, int dtype, int device);

/* Pinned host memory helpers for callers that want the fast NumPy path.     */
void *adrt_b200_host_alloc_pinned(size_t bytes);
void  adrt_b200_host_free_pinned(void *p);
/* 1 if `p` points into page-locked host memory known to CUDA (the host entry points then copy directly) */
int   adrt_b200_host_is_pinned(const void *p);

#ifdef __cplusplus
}
#endif
#endif /* ADRT_B200_H */
