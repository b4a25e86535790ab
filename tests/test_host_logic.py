"""CPU: the reference's argument-validation contract (exception type + message)
at the native boundary and in the Python wrappers, and the pure-NumPy utilities.
Cites /root/reference/tests/*.py for the behaviours being mirrored."""
import numpy as np
import pytest

import adrt_b200 as adrt
from adrt_b200 import _adrt_cdefs as cd
from adrt_b200 import _wrappers
from helpers import DTYPES, bytes_equal, make_sino


# ---- native boundary: strict layout (reference tests/test_adrt.py:135-166) ----
def test_cdefs_rejects_non_arrays():
    for fn in (cd.adrt, cd.bdrt, cd.iadrt, cd.interp_to_cart, cd.press_fmg_restriction,
               cd.press_fmg_prolongation, cd.press_fmg_highpass):
        with pytest.raises(TypeError, match="array must be a NumPy array or compatible subclass"):
            fn([[1.0, 2.0], [3.0, 4.0]])
        with pytest.raises(TypeError, match="array must be a NumPy array or compatible subclass"):
            fn(None)


def test_cdefs_rejects_bad_layouts():
    a = np.zeros((8, 8), dtype=np.float32)
    msg = "array must be C-order, contiguous, aligned, and native byte order"
    with pytest.raises(ValueError, match=msg):
        cd.adrt(np.asfortranarray(np.zeros((8, 8), dtype=np.float32) + np.arange(8)))
    with pytest.raises(ValueError, match=msg):
        cd.adrt(np.zeros((16, 16), dtype=np.float32)[::2, ::2])
    with pytest.raises(ValueError, match=msg):
        cd.adrt(a.astype(a.dtype.newbyteorder("S")))
    raw = np.zeros(8 * 8 * 4 + 1, dtype=np.uint8)
    unaligned = raw[1:].view(np.float32).reshape(8, 8)
    assert not unaligned.flags.aligned
    with pytest.raises(ValueError, match=msg):
        cd.adrt(unaligned)


def test_cdefs_dimension_messages():
    with pytest.raises(ValueError, match="array must have between 2 and 3 dimensions, but had 1"):
        cd.adrt(np.zeros(8, dtype=np.float32))
    with pytest.raises(ValueError, match="array must have between 2 and 3 dimensions, but had 4"):
        cd.adrt(np.zeros((1, 1, 8, 8), dtype=np.float32))
    with pytest.raises(ValueError, match="array must have between 3 and 4 dimensions, but had 2"):
        cd.bdrt(np.zeros((15, 8), dtype=np.float32))
    with pytest.raises(ValueError, match="all array dimensions must be nonzero, but found zero in dimension 0"):
        cd.adrt(np.zeros((0, 8, 8), dtype=np.float32))
    with pytest.raises(ValueError, match="found zero in dimension 1"):
        cd.adrt(np.zeros((3, 0, 8), dtype=np.float32))


@pytest.mark.parametrize("shape", [(8, 4), (6, 6), (3, 12, 12), (5, 8, 16)])
def test_adrt_shape_message(shape):
    with pytest.raises(ValueError, match="array must be square with a power of two shape"):
        cd.adrt(np.zeros(shape, dtype=np.float32))


def test_other_shape_messages():
    bad = np.zeros((4, 14, 8), dtype=np.float32)
    with pytest.raises(ValueError, match="array must have valid shape for ADRT, use adrt.core.adrt_init"):
        cd.adrt_step(bad, 0)
    for fn in (cd.bdrt, cd.iadrt, cd.interp_to_cart, cd.press_fmg_restriction):
        with pytest.raises(ValueError, match="array must have a valid ADRT output shape"):
            fn(bad)
    with pytest.raises(ValueError, match="array must have a valid ADRT output shape"):
        cd.bdrt_step(bad, 0)
    with pytest.raises(ValueError, match="array must have a valid ADRT output shape"):
        cd.bdrt(np.zeros((3, 15, 8), dtype=np.float32))
    # n = 1 is a valid ADRT shape but not for interp / restriction
    one = np.zeros((4, 1, 1), dtype=np.float32)
    with pytest.raises(ValueError, match="array must have a valid ADRT output shape"):
        cd.interp_to_cart(one)
    with pytest.raises(ValueError, match="array must have a valid ADRT output shape"):
        cd.press_fmg_restriction(one)
    with pytest.raises(ValueError, match="array is too small to high-pass filter"):
        cd.press_fmg_highpass(np.zeros((1, 8), dtype=np.float32))
    with pytest.raises(ValueError, match="array is too small to high-pass filter"):
        cd.press_fmg_highpass(np.zeros((3, 8, 1), dtype=np.float32))


def test_step_argument_contract():
    # reference tests/test_adrt_step.py:41-72
    a = np.zeros((4, 15, 8), dtype=np.float32)
    for bad in (-1, 3, 100):
        with pytest.raises(ValueError, match=f"step {bad} is out of range for array's shape, use adrt.core.num_iters"):
            cd.adrt_step(a, bad)
        with pytest.raises(ValueError, match=f"step {bad} is out of range"):
            cd.bdrt_step(a, bad)
    with pytest.raises(TypeError):
        cd.adrt_step(a, 1.0)
    with pytest.raises(TypeError):
        cd.adrt_step(a, "1")
    with pytest.raises(OverflowError, match="Python int too large to convert to C int"):
        cd.adrt_step(a, 2**40)
    with pytest.raises(TypeError):
        cd.adrt_step(a)
    with pytest.raises(TypeError):
        adrt.core.adrt_step(a, 1.5)
    one = np.zeros((4, 1, 1), dtype=np.float32)
    with pytest.raises(ValueError, match="step 0 is out of range"):
        cd.adrt_step(one, 0)


@pytest.mark.parametrize("dtype", [np.int32, np.float16, np.complex64, np.bool_])
def test_unsupported_dtype(dtype):
    with pytest.raises(TypeError, match="unsupported array dtype"):
        cd.adrt(np.zeros((8, 8), dtype=dtype))
    with pytest.raises(TypeError, match="unsupported array dtype"):
        cd.bdrt(np.zeros((4, 15, 8), dtype=dtype))
    # shape errors win over dtype errors (dtype switch is last, py.cpp:298-340)
    with pytest.raises(ValueError, match="power of two"):
        cd.adrt(np.zeros((8, 7), dtype=dtype))


# ---- public wrappers -------------------------------------------------------------
def test_wrapper_type_error():
    with pytest.raises(TypeError, match="array must be numpy.ndarray, but got list"):
        adrt.adrt([[1.0]])
    with pytest.raises(TypeError, match="array must be numpy.ndarray, but got NoneType"):
        adrt.bdrt(None)
    with pytest.raises(TypeError, match="array must be numpy.ndarray, but got int"):
        adrt.core.adrt_init(3)


def test_normalize_array():
    f = np.asfortranarray(np.arange(64, dtype=np.float32).reshape(8, 8))
    n = _wrappers._normalize_array(f)
    assert n.flags.c_contiguous and np.array_equal(n, f)
    s = np.arange(64, dtype=np.float32).reshape(8, 8).astype(">f4")
    n = _wrappers._normalize_array(s)
    assert n.dtype.isnative and np.array_equal(n, s)
    c = np.zeros((8, 8), dtype=np.float64)
    assert _wrappers._normalize_array(c) is c or np.shares_memory(_wrappers._normalize_array(c), c)


def test_num_iters():
    # reference tests/test_num_iters.py
    ni = adrt.core.num_iters
    assert [ni(i) for i in (0, 1, 2, 3, 4, 5, 8, 9, 1023, 1024, 1025)] == [0, 0, 1, 2, 2, 3, 3, 4, 10, 10, 11]
    assert ni(np.int32(16)) == 4
    assert ni(2**70) == 70
    with pytest.raises(ValueError, match="non-negative value required for iteration count, but got -1"):
        ni(-1)
    with pytest.raises(TypeError):
        ni(2.0)


def test_adrt_init_matches_golden(golden_small):
    for dn in DTYPES:
        for n in (1, 2, 4, 8, 16):
            for B in (0, 2):
                tag = f"{dn}_n{n}_b{B}"
                assert bytes_equal(adrt.core.adrt_init(golden_small[f"x_{tag}"]), golden_small[f"init_{tag}"])
    # any dtype is accepted (reference tests/test_adrt_init.py:40-49)
    xi = np.arange(16, dtype=np.int16).reshape(4, 4)
    out = adrt.core.adrt_init(xi)
    assert out.dtype == np.int16 and out.shape == (4, 7, 4)
    assert np.array_equal(out[2, :4], xi) and not out[:, 4:].any()
    with pytest.raises(ValueError, match="array must be square with a power of two shape"):
        adrt.core.adrt_init(np.zeros((6, 6)))
    with pytest.raises(ValueError, match="between 2 and 3 dimensions, but had 4"):
        adrt.core.adrt_init(np.zeros((1, 1, 4, 4)))


def test_truncate_stitch_match_golden(golden_small):
    for dn in DTYPES:
        for n in (1, 2, 4, 8, 16):
            for B in (0, 2):
                tag = f"{dn}_n{n}_b{B}"
                s = golden_small[f"s_{tag}"]
                assert bytes_equal(adrt.utils.truncate(s), golden_small[f"trunc_{tag}"])
                st = adrt.utils.stitch_adrt(s)
                assert bytes_equal(st, golden_small[f"stitch_{tag}"])
                assert bytes_equal(adrt.utils.stitch_adrt(s, remove_repeated=True), golden_small[f"stitchrr_{tag}"])
                assert bytes_equal(np.ascontiguousarray(adrt.utils.unstitch_adrt(st)), s)


def test_truncate_is_inverse_of_init():
    x = np.arange(3 * 8 * 8, dtype=np.float64).reshape(3, 8, 8)
    t = adrt.utils.truncate(adrt.core.adrt_init(x))
    for q in range(4):
        assert np.array_equal(t[:, q], x)


def test_utils_shape_errors():
    with pytest.raises(ValueError, match="unsuitable shape for ADRT output processing"):
        adrt.utils.truncate(np.zeros((4, 14, 8)))
    with pytest.raises(ValueError, match="unsuitable shape for ADRT output processing"):
        adrt.utils.stitch_adrt(np.zeros((3, 15, 8)))
    with pytest.raises(ValueError, match="unsuitable shape for ADRT unstitching"):
        adrt.utils.unstitch_adrt(np.zeros((21, 32)))
    # non power-of-two n is fine for these pure layout helpers
    s = make_sino(1, (2, 3, 4, 11, 6), np.float32)
    assert adrt.utils.stitch_adrt(s).shape == (2, 3, 16, 24)
    assert adrt.utils.truncate(s).shape == (2, 3, 4, 6, 6)
    assert bytes_equal(np.ascontiguousarray(adrt.utils.unstitch_adrt(adrt.utils.stitch_adrt(s))), s)


def test_coord_functions():
    c = adrt.utils.coord_adrt(8)
    assert c.offset.shape == (4, 15, 8) and c.angle.shape == (4, 1, 8)
    assert c.offset.dtype == np.float64
    assert np.allclose(c.angle[0, 0, 0], -np.pi / 2) and np.allclose(c.angle[3, 0, 0], np.pi / 2)
    assert np.allclose(c.angle[1, 0, -1], -np.pi / 4) and np.allclose(c.angle[2, 0, -1], np.pi / 4)
    with pytest.raises(ValueError, match="invalid Radon domain size 1, must be at least 2"):
        adrt.utils.coord_adrt(1)
    with pytest.raises(ValueError, match="invalid Radon domain size 6, must be a power of two"):
        adrt.utils.coord_adrt(6)
    th = np.linspace(-np.pi / 2, np.pi / 2, 9)
    t = np.zeros(9)
    idx = adrt.utils.coord_cart_to_adrt(th, t, 8)
    assert idx.quadrant.dtype == np.uint8 and idx.height.dtype == np.int64 and idx.slope.dtype == np.uint64
    assert idx.quadrant.min() >= 0 and idx.quadrant.max() <= 3
    with pytest.raises(ValueError, match="mismatched shapes for theta and t"):
        adrt.utils.coord_cart_to_adrt(th, t[:3], 8)


def test_iadrt_fmg_argument_errors():
    with pytest.raises(ValueError, match="batch dimension not supported for iadrt_fmg, got 4 dimensions"):
        adrt.iadrt_fmg(np.zeros((2, 4, 15, 8), dtype=np.float32))
    with pytest.raises(ValueError, match="must allow at least one iteration, but specified 0"):
        adrt.iadrt_fmg(np.zeros((4, 15, 8), dtype=np.float32), max_iters=0)


def test_threading_enabled_is_bool():
    assert isinstance(adrt.core.threading_enabled(), bool)
    assert isinstance(cd.OPENMP_ENABLED, bool)


# ---- round 2: out= aliasing, exact arity, the `adrt` import name -----------------
def test_out_must_not_overlap_input():
    # no transform runs in place: rejected before any device work
    y = np.zeros((4, 15, 8), dtype=np.float32)
    for fn in (cd.bdrt, cd.iadrt):
        with pytest.raises(ValueError, match="out must not overlap the input array"):
            fn(y, out=y)
    with pytest.raises(ValueError, match="out must not overlap"):
        cd.bdrt_step(y, 0, out=y)
    big = np.zeros(2 * y.size, dtype=np.float32)
    with pytest.raises(ValueError, match="out must not overlap"):
        cd.bdrt(big[: y.size].reshape(y.shape), out=big[4: 4 + y.size].reshape(y.shape))


def test_step_functions_check_arity_like_fastcall():
    # adrt_cdefs_py.cpp:200-211; reference tests/test_adrt_step.py:41-49
    arr = np.zeros((4, 31, 16), dtype=np.float32)
    for name in ("adrt_step", "bdrt_step"):
        fn = getattr(cd, name)
        with pytest.raises(TypeError, match=rf"{name} expected 2 arguments, got 1"):
            fn(arr)
        with pytest.raises(TypeError, match=rf"{name} expected 2 arguments, got 3"):
            fn(arr, 0, 0)
        with pytest.raises(TypeError, match="int"):
            fn(arr, [])


def test_adrt_import_name_is_the_engine():
    import adrt as ref_name
    import adrt.core
    import adrt.utils
    import adrt_b200

    assert ref_name.core is adrt_b200.core and ref_name.utils is adrt_b200.utils
    assert ref_name._adrt_cdefs is adrt_b200._adrt_cdefs and ref_name._wrappers is adrt_b200._wrappers
    assert ref_name.adrt is adrt_b200.adrt and ref_name.iadrt_fmg is adrt_b200.iadrt_fmg
    assert adrt.core.iadrt_fmg_iter is adrt_b200.core.iadrt_fmg_iter
    assert isinstance(ref_name._adrt_cdefs.OPENMP_ENABLED, bool)
    assert ref_name.core.threading_enabled() == ref_name._adrt_cdefs.OPENMP_ENABLED
    assert sorted(ref_name.__all__) == sorted(["adrt", "iadrt", "bdrt", "iadrt_fmg", "utils", "core"])


def test_tensor_subclass_and_dlpack_are_recognised():
    import torch

    class MyTensor(torch.Tensor):
        pass

    t = torch.zeros(2, 2).as_subclass(MyTensor)
    assert cd._is_torch_tensor(t) and cd._is_torch_tensor(torch.nn.Parameter(torch.zeros(1)))
    assert not cd._is_torch_tensor(np.zeros(2)) and not cd._is_torch_tensor(None)
    # CPU tensors are not arrays this engine accepts (same TypeError as any non-ndarray)
    with pytest.raises(TypeError, match="must be numpy.ndarray"):
        adrt.adrt(torch.zeros(4, 4))


def test_size_limit_is_a_value_error():
    # n > 16384 used to surface as a RuntimeError from the C ABI; now a ValueError raised before
    # anything is allocated (a real array of that size would be 34 GB, so the dispatcher is driven directly)
    fake = cd._Arr(np.zeros(1, dtype=np.float32), (4, 2 * 32768 - 1, 32768), np.dtype(np.float32), False)
    with pytest.raises(ValueError, match="image side 32768 exceeds the 16384"):
        cd._run("bdrt", fake, fake.shape, (1, 32768), workspace="bdrt")
    assert cd.MAX_SIDE == 16384
