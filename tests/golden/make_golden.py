"""Generate the committed golden fixtures from the UNMODIFIED reference.

Run in the authoring container only (needs /root/reference and
``make -C oracle ref``):

    python tests/golden/make_golden.py

Outputs (committed):
  tests/golden/small.npz   full input/output arrays for small sizes
  tests/golden/hashes.json sha256 of reference outputs on seeded inputs at
                           larger sizes (inputs are regenerated from the seed)

Inputs are drawn with ``np.random.default_rng(seed)`` (PCG64; stream is stable
across NumPy versions) and have negative zeros sprinkled in so that the
copy-vs-add rule of the reference (adrt_cdefs_adrt.hpp:78-81,
adrt_cdefs_bdrt.hpp:96-109) is observable in the bytes.
"""
import hashlib
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
sys.dont_write_bytecode = True

from oracle import ref_loader  # noqa: E402

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


def main():
    ref = ref_loader.load_ref_package()
    small = {}
    for dn, dt in DTYPES.items():
        for n in (1, 2, 4, 8, 16):
            for B in (0, 2):
                tag = f"{dn}_n{n}_b{B}"
                ishape = (n, n) if B == 0 else (B, n, n)
                x = make_image(100 + n + B, ishape, dt)
                y = ref.adrt(x)
                s = make_sino(200 + n + B, y.shape, dt)
                small[f"x_{tag}"] = x
                small[f"s_{tag}"] = s
                small[f"adrt_{tag}"] = y
                small[f"init_{tag}"] = ref.core.adrt_init(x)
                small[f"bdrt_{tag}"] = ref.bdrt(s)
                small[f"iadrt_{tag}"] = ref.iadrt(s)
                K = ref.core.num_iters(n)
                for i in range(K):
                    small[f"adrtstep{i}_{tag}"] = ref.core.adrt_step(s, i)
                    small[f"bdrtstep{i}_{tag}"] = ref.core.bdrt_step(s, i)
                small[f"prol_{tag}"] = ref._wrappers._press_fmg_prolongation(x)
                small[f"trunc_{tag}"] = ref.utils.truncate(s)
                small[f"stitch_{tag}"] = ref.utils.stitch_adrt(s)
                small[f"stitchrr_{tag}"] = ref.utils.stitch_adrt(s, remove_repeated=True)
                if n >= 2:
                    small[f"restr_{tag}"] = ref._wrappers._press_fmg_restriction(s)
                    small[f"highpass_{tag}"] = ref._wrappers._press_fmg_highpass(x)
                    small[f"interp_{tag}"] = ref.utils.interp_to_cart(s)
                    small[f"fmgstep_{tag}"] = ref.core.iadrt_fmg_step(s)
                if n >= 2 and B == 0:
                    a = ref.adrt(make_image(300 + n, (n, n), dt) * 0 + np.add.outer(
                        np.arange(n), np.arange(n)).astype(dt))
                    small[f"fmgin_{tag}"] = a
                    small[f"fmg3_{tag}"] = ref.iadrt_fmg(a, max_iters=3)
    np.savez_compressed(os.path.join(HERE, "small.npz"), **small)

    hashes = {}
    for dn, dt in DTYPES.items():
        for n, B in ((32, 3), (64, 3), (128, 2), (256, 1), (512, 1)):
            tag = f"{dn}_n{n}_b{B}"
            x = make_image(1000 + n, (B, n, n), dt)
            y = ref.adrt(x)
            s = make_sino(2000 + n, y.shape, dt)
            K = ref.core.num_iters(n)
            hashes[tag] = {
                "adrt": sha(y),
                "bdrt_of_adrt": sha(ref.bdrt(y)),
                "bdrt": sha(ref.bdrt(s)),
                "iadrt": sha(ref.iadrt(s)),
                "adrt_step": [sha(ref.core.adrt_step(s, i)) for i in range(K)],
                "bdrt_step": [sha(ref.core.bdrt_step(s, i)) for i in range(K)],
                "interp": sha(ref.utils.interp_to_cart(s)),
                "restr": sha(ref._wrappers._press_fmg_restriction(s)),
                "highpass": sha(ref._wrappers._press_fmg_highpass(x)),
                "prol": sha(ref._wrappers._press_fmg_prolongation(x)),
                "fmgstep": sha(ref.core.iadrt_fmg_step(s)),
            }
    # BASELINE.json configs[0]: 256x256 float32 uniform, forward only
    x = np.random.default_rng(0).random((256, 256), dtype=np.float32)
    hashes["config0_adrt_256_f32_uniform_seed0"] = sha(ref.adrt(x))
    with open(os.path.join(HERE, "hashes.json"), "w") as f:
        json.dump(hashes, f, indent=1, sort_keys=True)
    print("wrote", len(small), "arrays and", len(hashes), "hash groups")


if __name__ == "__main__":
    main()
