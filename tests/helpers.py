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
