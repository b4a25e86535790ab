"""Drop-in for the reference's native module ``adrt._adrt_cdefs``.

Same nine functions, same positional-only signatures, same validation order
and the same exception types / messages as
/root/reference/src/adrt/adrt_cdefs_py.cpp:275-850 (module table :852-863),
but the compute call goes to hand-written sm_100a CUDA through the C ABI in
``include/adrt_b200.h`` instead of the C++/OpenMP templates.

Two kinds of array are accepted:

* ``numpy.ndarray`` -- the reference contract (C-order, aligned, native byte
  order, float32/float64).  Data moves host -> B200 -> host inside the call
  (``adrt_b200_host_*``); the result is a fresh writable ndarray.
* CUDA ``torch.Tensor`` (an extension; also anything exposing ``__dlpack__``
  on a CUDA device, via ``torch.from_dlpack``) -- zero-copy, runs on the
  tensor's device on torch's current stream, returns a tensor.

Extension keyword ``out=`` lets callers supply the (e.g. pinned) result buffer.
There is no CPU implementation here and none is ever substituted.
"""
from __future__ import annotations

import operator
import os

import numpy as np

from . import _lib

# The reference exports whether its core was built with OpenMP
# (adrt_cdefs_py.cpp:896).  The B200 engine always runs massively threaded.
OPENMP_ENABLED = True

_INT_MAX = 2**31 - 1
_INT_MIN = -(2**31)
_SIZE_BITS = 64
_MAX_SIZE = 1 << (_SIZE_BITS - 1)  # adrt_cdefs_common.cpp:62

# Largest image side n the kernels index (32-bit in-plane offsets, csrc/common.cuh kMaxN): a float32
# (4, 2n-1, n) array of that size is 8.6 GB.  Larger power-of-two sides raise ValueError here.
MAX_SIDE = 16384

_device = None


def set_device(index: int) -> None:
    """Select the CUDA ordinal used by the NumPy path (default: $LOCAL_RANK or 0)."""
    global _device
    _device = int(index)


def get_device() -> int:
    global _device
    if _device is None:
        _device = int(os.environ.get("ADRT_B200_DEVICE", os.environ.get("LOCAL_RANK", "0")))
    return _device


# ---------------------------------------------------------------------------
# argument handling
# ---------------------------------------------------------------------------

def _is_torch_tensor(a) -> bool:
    # torch is imported lazily (the NumPy path does not need it): only look it up
    # when the object's class hierarchy mentions torch at all (covers Tensor subclasses)
    if not any(c.__module__.split(".")[0] == "torch" for c in type(a).__mro__):
        return False
    import torch

    return isinstance(a, torch.Tensor)


class _Arr:
    """Normalised view of an input: shape, dtype code and how to get pointers."""

    __slots__ = ("obj", "shape", "ndim", "np_dtype", "is_torch")

    def __init__(self, obj, shape, np_dtype, is_torch):
        self.obj = obj
        self.shape = tuple(int(s) for s in shape)
        self.ndim = len(self.shape)
        self.np_dtype = np_dtype
        self.is_torch = is_torch


def _extract_array(a) -> _Arr:
    # adrt_cdefs_py.cpp:77-90 (extract_array)
    if isinstance(a, np.ndarray):
        f = a.flags
        if not (f.c_contiguous and f.aligned and a.dtype.isnative):
            raise ValueError("array must be C-order, contiguous, aligned, and native byte order")
        return _Arr(a, a.shape, a.dtype, False)
    if not _is_torch_tensor(a) and hasattr(a, "__dlpack__") and hasattr(a, "__dlpack_device__"):
        dev_type = int(a.__dlpack_device__()[0])
        if dev_type == 2:  # kDLCUDA
            import torch

            a = torch.from_dlpack(a)
    if _is_torch_tensor(a):
        import torch

        if not a.is_cuda:
            raise TypeError("array must be a NumPy array or compatible subclass")
        if not a.is_contiguous():
            raise ValueError("array must be C-order, contiguous, aligned, and native byte order")
        a = a.detach()
        np_dtype = {torch.float32: np.dtype(np.float32), torch.float64: np.dtype(np.float64)}.get(a.dtype)
        if np_dtype is None:
            np_dtype = str(a.dtype)
        return _Arr(a, a.shape, np_dtype, True)
    raise TypeError("array must be a NumPy array or compatible subclass")


def _array_shape(arr: _Arr, min_dim: int, max_dim: int):
    # adrt_cdefs_py.cpp:121-154 (array_shape): left-pad to max_dim
    if arr.ndim < min_dim or arr.ndim > max_dim:
        raise ValueError(
            f"array must have between {min_dim} and {max_dim} dimensions, but had {arr.ndim}"
        )
    for i, s in enumerate(arr.shape):
        if s <= 0:
            raise ValueError(f"all array dimensions must be nonzero, but found zero in dimension {i}")
    return (1,) * (max_dim - arr.ndim) + arr.shape


def _extract_int(v) -> int:
    # adrt_cdefs_py.cpp:100-119 (extract_int): PyLong_AsLong then range check
    if isinstance(v, float):
        raise TypeError("'float' object cannot be interpreted as an integer")
    try:
        val = operator.index(v)
    except TypeError:
        raise TypeError(f"'{type(v).__name__}' object cannot be interpreted as an integer") from None
    if val > _INT_MAX or val < _INT_MIN:
        raise OverflowError("Python int too large to convert to C int")
    return val


def _dtype_code(arr: _Arr) -> int:
    # adrt_cdefs_py.cpp:264-269 (report_unsupported_dtype)
    if arr.np_dtype == np.float32:
        return _lib.F32
    if arr.np_dtype == np.float64:
        return _lib.F64
    raise TypeError(f"unsupported array dtype {arr.np_dtype}")


def _is_pow2(n: int) -> bool:
    return n > 0 and (n & (n - 1)) == 0


def _num_iters(n: int) -> int:
    return n.bit_length() - (1 if _is_pow2(n) else 0)


def _is_adrt_output_shape(shape4) -> bool:
    # adrt_cdefs_common.cpp:185-192 (adrt_step_is_valid_shape)
    _, q, d, n = shape4
    return q == 4 and n <= _MAX_SIZE and d == 2 * n - 1 and _is_pow2(n)


def _result_shape(arr: _Arr, virtual_shape, drop: int = 0):
    # adrt_cdefs_py.cpp:156-186 (new_array): keep the caller's ndim (minus `drop`)
    nd = arr.ndim - drop
    return tuple(virtual_shape[len(virtual_shape) - nd:])


def _tensor_span(t):
    """[first, last) byte addresses a contiguous tensor occupies."""
    start = t.data_ptr()
    return start, start + t.numel() * t.element_size()


def _empty_like(arr: _Arr, shape, out, alias_ok=False):
    # No transform runs in place (CTAs read tiles of the input while others write the
    # output), so an `out` that overlaps the input is rejected instead of racing.
    if arr.is_torch:
        import torch

        if out is not None:
            if not (_is_torch_tensor(out) and out.is_cuda and out.is_contiguous()
                    and tuple(out.shape) == tuple(shape) and out.dtype == arr.obj.dtype
                    and out.device == arr.obj.device):
                raise ValueError("out must be a contiguous CUDA tensor of the result shape, dtype and device")
            (a0, a1), (b0, b1) = _tensor_span(arr.obj), _tensor_span(out)
            if not alias_ok and a0 < b1 and b0 < a1:
                raise ValueError("out must not overlap the input array")
            return out
        return torch.empty(shape, dtype=arr.obj.dtype, device=arr.obj.device)
    if out is not None:
        if not (isinstance(out, np.ndarray) and out.flags.c_contiguous and out.flags.aligned
                and out.flags.writeable and out.shape == tuple(shape) and out.dtype == arr.np_dtype):
            raise ValueError("out must be a writable C-contiguous ndarray of the result shape and dtype")
        if not alias_ok and np.may_share_memory(out, arr.obj):
            raise ValueError("out must not overlap the input array")
        return out
    # large results: fresh arrays in page-locked memory the GPU writes directly (_pinned.py)
    from . import _pinned

    fresh = None
    try:
        fresh = _pinned.empty(shape, arr.np_dtype)
    except Exception:   # no library / no device yet: the ordinary array; the call itself reports the problem
        fresh = None
    return fresh if fresh is not None else np.empty(shape, dtype=arr.np_dtype)


# The fused transforms (adrt, bdrt and the helpers built on them) use 16/32-byte vector accesses and
# the C ABI asks for 32-byte aligned device pointers.  Allocator-owned tensors always are; a view that
# starts mid-allocation is staged through fresh storage.  (The step kernels, iadrt and the FMG
# operators pick scalar variants for such pointers themselves.)
_VEC_OPS = frozenset({"adrt", "bdrt"})


def _vec_in(t):
    return t if t.data_ptr() % 32 == 0 else t.clone()


def _vec_out(ret):
    """(tensor to hand to the kernel, copy-back needed)."""
    if ret.data_ptr() % 32 == 0:
        return ret, False
    import torch

    return torch.empty_like(ret), True


def _run(name, arr: _Arr, out_shape, dims, out=None, step=None, workspace=None, extra=()):
    """Dispatch to adrt_b200_host_<name> (NumPy) or adrt_b200_<name> (CUDA tensor).

    ``dims`` are the int64 size arguments of the C entry point (B, n) or (B, h, w).
    ``workspace`` names the adrt_b200_<x>_workspace_bytes query, if the op has one.
    """
    lib = _lib.load()
    code = _dtype_code(arr)
    if len(dims) == 2 and dims[1] > MAX_SIDE:
        # documented limit of this engine (32-bit in-plane indices: (2n-1) n < 2^31); the reference has none
        raise ValueError(f"array is too big: image side {dims[1]} exceeds the {MAX_SIDE} this engine supports")
    ret = _empty_like(arr, out_shape, out)
    _lib.require_device()
    step_args = () if step is None else (step,)
    if not arr.is_torch:
        fn = getattr(lib, f"adrt_b200_host_{name}")
        rc = fn(arr.obj.ctypes.data, ret.ctypes.data, *dims, *step_args, code, get_device())
        _lib.check(rc, name)
        return ret
    import torch

    t = arr.obj
    with torch.cuda.device(t.device):
        stream = torch.cuda.current_stream().cuda_stream
        fn = getattr(lib, f"adrt_b200_{name}")
        dst, copy_back = ret, False
        if name in _VEC_OPS:
            t = _vec_in(t)
            dst, copy_back = _vec_out(ret)
        if workspace is not None:
            nbytes = getattr(lib, f"adrt_b200_{workspace}_workspace_bytes")(*dims, code)
            ws = torch.empty(max(int(nbytes), 1), dtype=torch.uint8, device=t.device)
            # `ws` is freed to torch's caching allocator on return; the allocator only
            # re-issues it to work queued later on this same stream, so that is safe.
            rc = fn(t.data_ptr(), dst.data_ptr(), *dims, *step_args, code, ws.data_ptr(), int(nbytes), stream)
        else:
            rc = fn(t.data_ptr(), dst.data_ptr(), *dims, *step_args, *extra, code, stream)
        _lib.check(rc, name)
        if copy_back:
            ret.copy_(dst)
    return ret


# ---------------------------------------------------------------------------
# the nine functions of adrt._adrt_cdefs
# ---------------------------------------------------------------------------

def adrt(a, /, *, out=None):
    """adrt_cdefs_py.cpp:275-341 -> (B?,n,n) to (B?,4,2n-1,n)."""
    arr = _extract_array(a)
    shape = _array_shape(arr, 2, 3)
    b, r, c = shape
    if not (r == c and c <= _MAX_SIZE and _is_pow2(c)):
        raise ValueError("array must be square with a power of two shape")
    res = _result_shape(arr, (b, 4, 2 * c - 1, c), drop=-1)
    return _run("adrt", arr, res, (b, c), out=out, workspace="adrt")


def _two_args(name, args):
    # METH_FASTCALL with an exact-arity check (adrt_cdefs_py.cpp:200-211)
    if len(args) != 2:
        raise TypeError(f"{name} expected 2 arguments, got {len(args)}")
    return args


def adrt_step(*args, out=None):
    """adrt_cdefs_py.cpp:343-412; ``adrt_step(a, step, /)``."""
    a, step = _two_args("adrt_step", args)
    arr = _extract_array(a)
    shape = _array_shape(arr, 3, 4)
    if not _is_adrt_output_shape(shape):
        raise ValueError("array must have valid shape for ADRT, use adrt.core.adrt_init")
    it = _extract_int(step)
    if not (0 <= it < _num_iters(shape[3])):
        raise ValueError(f"step {it} is out of range for array's shape, use adrt.core.num_iters")
    return _run("adrt_step", arr, arr.shape, (shape[0], shape[3]), out=out, step=it)


def iadrt(a, /, *, out=None):
    """adrt_cdefs_py.cpp:414-480."""
    arr = _extract_array(a)
    shape = _array_shape(arr, 3, 4)
    if not _is_adrt_output_shape(shape):
        raise ValueError("array must have a valid ADRT output shape")
    return _run("iadrt", arr, arr.shape, (shape[0], shape[3]), out=out, workspace="iadrt")


def bdrt(a, /, *, out=None):
    """adrt_cdefs_py.cpp:482-548."""
    arr = _extract_array(a)
    shape = _array_shape(arr, 3, 4)
    if not _is_adrt_output_shape(shape):
        raise ValueError("array must have a valid ADRT output shape")
    return _run("bdrt", arr, arr.shape, (shape[0], shape[3]), out=out, workspace="bdrt")


def bdrt_step(*args, out=None):
    """adrt_cdefs_py.cpp:550-619; ``bdrt_step(a, step, /)``."""
    a, step = _two_args("bdrt_step", args)
    arr = _extract_array(a)
    shape = _array_shape(arr, 3, 4)
    if not _is_adrt_output_shape(shape):
        raise ValueError("array must have a valid ADRT output shape")
    it = _extract_int(step)
    if not (0 <= it < _num_iters(shape[3])):
        raise ValueError(f"step {it} is out of range for array's shape, use adrt.core.num_iters")
    return _run("bdrt_step", arr, arr.shape, (shape[0], shape[3]), out=out, step=it)


def interp_to_cart(a, /, *, out=None):
    """adrt_cdefs_py.cpp:621-682."""
    arr = _extract_array(a)
    shape = _array_shape(arr, 3, 4)
    if not (_is_adrt_output_shape(shape) and shape[3] > 1):
        raise ValueError("array must have a valid ADRT output shape")
    n = shape[3]
    if n > (1 << 24) // 4:  # interp_adrtcart_is_valid_float_index<float>
        raise ValueError("array is too big for interpolation index calculations")
    res = _result_shape(arr, (shape[0], n, 4 * n), drop=1)
    return _run("interp_to_cart", arr, res, (shape[0], n), out=out)


def press_fmg_restriction(a, /, *, out=None):
    """adrt_cdefs_py.cpp:684-741."""
    arr = _extract_array(a)
    shape = _array_shape(arr, 3, 4)
    b, q, d, n = shape
    if not (q == 4 and n <= _MAX_SIZE and d == 2 * n - 1 and n >= 2 and n % 2 == 0):
        raise ValueError("array must have a valid ADRT output shape")
    res = _result_shape(arr, (b, 4, n - 1, n // 2))
    return _run("fmg_restriction", arr, res, (b, n), out=out)


def press_fmg_prolongation(a, /, *, out=None):
    """adrt_cdefs_py.cpp:743-797."""
    arr = _extract_array(a)
    shape = _array_shape(arr, 2, 3)
    b, h, w = shape
    lim = (2**_SIZE_BITS - 1) // 2
    if not (h <= lim and w <= lim):
        raise ValueError("array is too large for prolongation operator")
    res = _result_shape(arr, (b, 2 * h, 2 * w))
    return _run("fmg_prolongation", arr, res, (b, h, w), out=out)


def press_fmg_highpass(a, /, *, out=None):
    """adrt_cdefs_py.cpp:799-850."""
    arr = _extract_array(a)
    shape = _array_shape(arr, 2, 3)
    b, h, w = shape
    if not (h >= 2 and w >= 2):
        raise ValueError("array is too small to high-pass filter")
    return _run("fmg_highpass", arr, arr.shape, (b, h, w), out=out)


# ---------------------------------------------------------------------------
# extensions used by the device-resident drivers (not in the reference module)
# ---------------------------------------------------------------------------

def adrt_init(a, /, *, out=None):
    """Device version of core.adrt_init (core.py:123-176) for float arrays."""
    arr = _extract_array(a)
    shape = _array_shape(arr, 2, 3)
    b, r, c = shape
    if not (r == c and _is_pow2(c)):
        raise ValueError("array must be square with a power of two shape")
    res = _result_shape(arr, (b, 4, 2 * c - 1, c), drop=-1)
    return _run("adrt_init", arr, res, (b, c), out=out)


def truncate_mean(a, divisor, /, *, out=None):
    """mean over quadrants of truncate(a) / divisor, NumPy summation order
    (core.py:329: ``np.mean(truncate(x) / (n - 1), axis=-3)``); CUDA tensors only."""
    arr = _extract_array(a)
    if not arr.is_torch:
        raise TypeError("truncate_mean is a device-only helper (CUDA tensors)")
    shape = _array_shape(arr, 3, 4)
    if not _is_adrt_output_shape(shape):
        raise ValueError("array must have a valid ADRT output shape")
    b, _, _, n = shape
    res = _result_shape(arr, (b, n, n), drop=1)
    lib = _lib.load()
    code = _dtype_code(arr)
    ret = _empty_like(arr, res, out)
    import torch

    with torch.cuda.device(arr.obj.device):
        rc = lib.adrt_b200_truncate_mean(arr.obj.data_ptr(), ret.data_ptr(), b, n, float(divisor), code,
                                         torch.cuda.current_stream().cuda_stream)
    _lib.check(rc, "truncate_mean")
    return ret


def fmg_step(a, /, *, out=None):
    """core.iadrt_fmg_step (core.py:265-331) for CUDA tensors ``(B?, 4, 2n-1, n)`` ->
    ``(B?, n, n)``: all multigrid levels in one native call."""
    arr = _extract_array(a)
    if not arr.is_torch:
        raise TypeError("fmg_step is a device-only helper (CUDA tensors)")
    shape = _array_shape(arr, 3, 4)
    if not _is_adrt_output_shape(shape):
        raise ValueError("array must have a valid ADRT output shape")
    b, _, _, n = shape
    res = _result_shape(arr, (b, n, n), drop=1)
    lib = _lib.load()
    code = _dtype_code(arr)
    ret = _empty_like(arr, res, out)
    import torch

    with torch.cuda.device(arr.obj.device):
        nbytes = int(lib.adrt_b200_fmg_step_workspace_bytes(b, n, code))
        ws = torch.empty(max(nbytes, 1), dtype=torch.uint8, device=arr.obj.device)
        src = _vec_in(arr.obj)
        dst, copy_back = _vec_out(ret)
        rc = lib.adrt_b200_fmg_step(src.data_ptr(), dst.data_ptr(), b, n, code, ws.data_ptr(), nbytes,
                                    torch.cuda.current_stream().cuda_stream)
        _lib.check(rc, "fmg_step")
        if copy_back:
            ret.copy_(dst)
    return ret


def truncate(a, /, *, out=None):
    """Device version of utils.truncate (utils.py:231-242); CUDA tensors only."""
    arr = _extract_array(a)
    if not arr.is_torch:
        raise TypeError("truncate is a device-only helper (CUDA tensors)")
    shape = _array_shape(arr, 3, 4)
    if not _is_adrt_output_shape(shape):
        raise ValueError("array must have a valid ADRT output shape")
    b, _, _, n = shape
    res = _result_shape(arr, (b, 4, n, n))
    lib = _lib.load()
    code = _dtype_code(arr)
    ret = _empty_like(arr, res, out)
    import torch

    with torch.cuda.device(arr.obj.device):
        rc = lib.adrt_b200_truncate(arr.obj.data_ptr(), ret.data_ptr(), b, n, code,
                                    torch.cuda.current_stream().cuda_stream)
    _lib.check(rc, "truncate")
    return ret


def _binary(name, a, b, out=None):
    """Elementwise ``a - b`` / ``a + b`` on CUDA tensors of identical shape/dtype."""
    import torch

    if not (_is_torch_tensor(a) and _is_torch_tensor(b) and a.is_cuda and b.is_cuda):
        raise TypeError(f"{name} is a device-only helper (CUDA tensors)")
    if a.shape != b.shape or a.dtype != b.dtype or a.device != b.device:
        raise ValueError(f"{name}: operands must match in shape, dtype and device")
    a, b = a.contiguous(), b.contiguous()
    arr = _extract_array(a)
    code = _dtype_code(arr)
    ret = _empty_like(arr, arr.shape, out, alias_ok=True)  # elementwise: in place is fine
    lib = _lib.load()
    with torch.cuda.device(a.device):
        rc = getattr(lib, f"adrt_b200_{name}")(a.data_ptr(), b.data_ptr(), ret.data_ptr(), a.numel(), code,
                                                torch.cuda.current_stream().cuda_stream)
    _lib.check(rc, name)
    return ret


def sub(a, b, /, *, out=None):
    return _binary("sub", a, b, out)


def add(a, b, /, *, out=None):
    return _binary("add", a, b, out)


def adrt_quadrants(a, q_first, q_count, /, *, out=None):
    """Quadrants ``q_first .. q_first+q_count-1`` of ``adrt(a)`` only:
    ``(B?, n, n)`` -> ``(B?, q_count, 2n-1, n)``.  CUDA tensors only; used to shard
    one large image over several GPUs (the four quadrants are independent)."""
    arr = _extract_array(a)
    if not arr.is_torch:
        raise TypeError("adrt_quadrants is a device-only helper (CUDA tensors)")
    shape = _array_shape(arr, 2, 3)
    b, r, c = shape
    if not (r == c and _is_pow2(c)):
        raise ValueError("array must be square with a power of two shape")
    q_first, q_count = operator.index(q_first), operator.index(q_count)
    if not (0 <= q_first and 1 <= q_count and q_first + q_count <= 4):
        raise ValueError(f"bad quadrant range {q_first}+{q_count}")
    res = _result_shape(arr, (b, q_count, 2 * c - 1, c), drop=-1)
    lib = _lib.load()
    code = _dtype_code(arr)
    _lib.require_device()
    ret = _empty_like(arr, res, out)
    import torch

    t = arr.obj
    with torch.cuda.device(t.device):
        nbytes = lib.adrt_b200_adrt_quadrants_workspace_bytes(b, c, code, q_count)
        ws = torch.empty(max(int(nbytes), 1), dtype=torch.uint8, device=t.device)
        t = _vec_in(t)
        dst, copy_back = _vec_out(ret)
        rc = lib.adrt_b200_adrt_quadrants(t.data_ptr(), dst.data_ptr(), b, c, code, q_first, q_count,
                                          ws.data_ptr(), int(nbytes), torch.cuda.current_stream().cuda_stream)
        _lib.check(rc, "adrt_quadrants")
        if copy_back:
            ret.copy_(dst)
    return ret


def bdrt_planes(a, /, *, out=None, rows=None):
    """Back-projection of independent ``(2n-1, n)`` planes: ``(..., 2n-1, n)`` -> same
    shape (any leading dims, e.g. a subset of quadrants).  CUDA tensors only.

    ``rows``: compute only offsets ``d < rows`` of every plane (the rest of the
    result is unspecified) -- ``rows=n`` is all that ``truncate`` reads."""
    if not (_is_torch_tensor(a) and a.is_cuda):
        raise TypeError("bdrt_planes is a device-only helper (CUDA tensors)")
    a = a.contiguous()
    n = a.shape[-1]
    if a.ndim < 2 or a.shape[-2] != 2 * n - 1 or not _is_pow2(n):
        raise ValueError("array must have a valid ADRT output shape")
    arr = _extract_array(a)
    code = _dtype_code(arr)
    planes = 1
    for s in a.shape[:-2]:
        planes *= int(s)
    if planes < 1:
        raise ValueError("all array dimensions must be nonzero, but found zero in dimension 0")
    lib = _lib.load()
    _lib.require_device()
    ret = _empty_like(arr, arr.shape, out)
    import torch

    with torch.cuda.device(a.device):
        nbytes = lib.adrt_b200_bdrt_planes_workspace_bytes(planes, n, code)
        ws = torch.empty(max(int(nbytes), 1), dtype=torch.uint8, device=a.device)
        a = _vec_in(a)
        dst, copy_back = _vec_out(ret)
        if rows is None:
            rc = lib.adrt_b200_bdrt_planes(a.data_ptr(), dst.data_ptr(), planes, n, code, ws.data_ptr(), int(nbytes),
                                           torch.cuda.current_stream().cuda_stream)
        else:
            rc = lib.adrt_b200_bdrt_rows(a.data_ptr(), dst.data_ptr(), planes, n, int(rows), code, ws.data_ptr(),
                                         int(nbytes), torch.cuda.current_stream().cuda_stream)
        _lib.check(rc, "bdrt_planes")
        if copy_back:
            ret.copy_(dst)
    return ret


def normal_operator(a, divisor=1.0, /, *, out=None):
    """``mean_q(truncate(bdrt(adrt(a))) / divisor)`` for CUDA tensors ``(B?, n, n)`` in one native
    call (``adrt_b200_normal_operator``): the operator ``A^T A`` of the CG recipe
    (docs/examples.cginverse.md:45-52).  The sinogram goes from adrt to bdrt as workspace rows and
    never takes the public layout; bit-identical to ``truncate_mean(bdrt(adrt(a)), divisor)``."""
    arr = _extract_array(a)
    if not arr.is_torch:
        raise TypeError("normal_operator is a device-only helper (CUDA tensors)")
    shape = _array_shape(arr, 2, 3)
    b, r, c = shape
    if not (r == c and _is_pow2(c)):
        raise ValueError("array must be square with a power of two shape")
    code = _dtype_code(arr)
    lib = _lib.load()
    _lib.require_device()
    ret = _empty_like(arr, arr.shape, out)
    import torch

    t = arr.obj
    with torch.cuda.device(t.device):
        nbytes = int(lib.adrt_b200_normal_operator_workspace_bytes(b, c, code))
        ws = torch.empty(max(nbytes, 1), dtype=torch.uint8, device=t.device)
        t = _vec_in(t)
        dst, copy_back = _vec_out(ret)
        rc = lib.adrt_b200_normal_operator(t.data_ptr(), dst.data_ptr(), b, c, float(divisor), code, ws.data_ptr(), nbytes,
                                           torch.cuda.current_stream().cuda_stream)
        _lib.check(rc, "normal_operator")
        if copy_back:
            ret.copy_(dst)
    return ret


def adrt_bdrt_rows(a, q_first, q_count, /, *, out=None):
    """Offsets ``d < n`` of ``bdrt(adrt(a))`` for quadrants ``q_first .. q_first+q_count-1``:
    ``(B?, n, n)`` -> ``(B?, q_count, 2n-1, n)`` (rows ``d >= n`` unspecified).  The building block
    of the quadrant-sharded normal operator; CUDA tensors only, ``n >= 2``."""
    arr = _extract_array(a)
    if not arr.is_torch:
        raise TypeError("adrt_bdrt_rows is a device-only helper (CUDA tensors)")
    shape = _array_shape(arr, 2, 3)
    b, r, c = shape
    if not (r == c and _is_pow2(c) and c >= 2):
        raise ValueError("array must be square with a power of two shape")
    q_first, q_count = operator.index(q_first), operator.index(q_count)
    if not (0 <= q_first and 1 <= q_count and q_first + q_count <= 4):
        raise ValueError(f"bad quadrant range {q_first}+{q_count}")
    res = _result_shape(arr, (b, q_count, 2 * c - 1, c), drop=-1)
    code = _dtype_code(arr)
    lib = _lib.load()
    _lib.require_device()
    ret = _empty_like(arr, res, out)
    import torch

    t = arr.obj
    with torch.cuda.device(t.device):
        nbytes = int(lib.adrt_b200_adrt_bdrt_rows_workspace_bytes(b, c, code, q_count))
        ws = torch.empty(max(nbytes, 1), dtype=torch.uint8, device=t.device)
        t = _vec_in(t)
        dst, copy_back = _vec_out(ret)
        rc = lib.adrt_b200_adrt_bdrt_rows(t.data_ptr(), dst.data_ptr(), b, c, code, q_first, q_count, ws.data_ptr(), nbytes,
                                          torch.cuda.current_stream().cuda_stream)
        _lib.check(rc, "adrt_bdrt_rows")
        if copy_back:
            ret.copy_(dst)
    return ret


def bdrt_truncate_mean(a, divisor, /, *, out=None):
    """``truncate_mean(bdrt(a), divisor)`` for CUDA tensors ``(B?, 4, 2n-1, n)``: the
    back-projection computes only the offsets ``d < n`` that ``truncate`` keeps
    (utils.py:231-242), bit-identical to the unfused composition."""
    if not (_is_torch_tensor(a) and a.is_cuda):
        raise TypeError("bdrt_truncate_mean is a device-only helper (CUDA tensors)")
    shape = _array_shape(_extract_array(a), 3, 4)
    if not _is_adrt_output_shape(shape):
        raise ValueError("array must have a valid ADRT output shape")
    return truncate_mean(bdrt_planes(a, rows=shape[-1]), divisor, out=out)
