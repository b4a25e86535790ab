"""Shared input generators (identical to tests/golden/make_golden.py)."""
import hashlib

import numpy as np

DTYPES = {"f32": np.float32, "f64": np.float64}


def make_image(seed, shape, dtype):
    rng = np.random.default_rng(seed)
    x = rng.standard_normal(shape).astype(dtype)
    x.flat[::5] = -0.0
    return x


def make_sino(seed, shape, dtype):
    rng = np.random.default_rng(seed)
    s = rng.standard_normal(shape).astype(dtype)
    s.flat[::7] = -0.0
    s[..., -1, :] = -0.0
    return s


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def bytes_equal(a, b):
    a = np.asarray(a)
    b = np.asarray(b)
    return a.shape == b.shape and a.dtype == b.dtype and a.tobytes() == b.tobytes()


def first_diff(a, b):
    """Human-readable description of the first differing element."""
    a = np.asarray(a)
    b = np.asarray(b)
    if a.shape != b.shape or a.dtype != b.dtype:
        return f"shape/dtype {a.shape}/{a.dtype} vs {b.shape}/{b.dtype}"
    av = a.view(np.uint32 if a.dtype == np.float32 else np.uint64)
    bv = b.view(av.dtype)
    bad = np.argwhere(av != bv)
    if len(bad) == 0:
        return "identical"
    i = tuple(bad[0])
    return f"{len(bad)} of {a.size} differ; first at {i}: got {a[i]!r} want {b[i]!r}"


# ---- the reference's multigrid pass driven by its own compiled core ---------------
def ref_truncate(a):
    """utils.py:231-242 (test-side NumPy twin of the reference's ``truncate``)."""
    n = a.shape[-1]
    return np.stack(
        [
            np.flip(a[..., 0, :n, :n], axis=-2).swapaxes(-1, -2),
            np.flip(a[..., 1, :n, :n], axis=-2),
            a[..., 2, :n, :n],
            np.flip(a[..., 3, :n, :n], axis=(-1, -2)).swapaxes(-1, -2),
        ],
        axis=-3,
    )


def ref_fmg_step(ref, a):
    """core.py:318-331 with every native operator taken from `ref`, the reference's own
    ``_adrt_cdefs`` module (oracle/_ref, multithreaded), and the glue in NumPy exactly as
    the reference spells it (``np.mean(truncate(.) / (n - 1), axis=-3)``).  Fast enough
    for BASELINE config 4's 4096^2 on the GPU box's host cores."""
    stack = []
    m = a.shape[-1]
    while m > 1:
        stack.append(a)
        a = ref.press_fmg_restriction(np.ascontiguousarray(a))
        m //= 2
    ret = np.ascontiguousarray(a[..., 0, :, :])
    n = 1
    while stack:
        n *= 2
        ret = ref.press_fmg_prolongation(ret)
        resid = ref.adrt(ret) - stack.pop()
        ret -= ref.press_fmg_highpass(
            np.ascontiguousarray(np.mean(ref_truncate(ref.bdrt(resid)) / (n - 1), axis=-3)))
    return ret
