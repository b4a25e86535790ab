"""CPU: pin the oracle (oracle/adrt_oracle.c + oracle/oracle.py) against the
golden vectors generated from the unmodified reference, and -- where the
compiled reference core is present (oracle/_ref) -- against the reference live."""
import numpy as np
import pytest

from helpers import DTYPES, bytes_equal, first_diff, make_image, make_sino, sha
from oracle import oracle as O
from oracle import ref_loader

CASES = [(dn, n, B) for dn in DTYPES for n in (1, 2, 4, 8, 16) for B in (0, 2)]


@pytest.mark.parametrize("dn,n,B", CASES)
def test_oracle_vs_golden_small(golden_small, dn, n, B):
    g = golden_small
    tag = f"{dn}_n{n}_b{B}"
    x, s = g[f"x_{tag}"], g[f"s_{tag}"]
    checks = {
        "adrt": (O.adrt(x), g[f"adrt_{tag}"]),
        "init": (O.adrt_init(x), g[f"init_{tag}"]),
        "bdrt": (O.bdrt(s), g[f"bdrt_{tag}"]),
        "iadrt": (O.iadrt(s), g[f"iadrt_{tag}"]),
        "prol": (O.press_fmg_prolongation(x), g[f"prol_{tag}"]),
        "trunc": (O.truncate(s), g[f"trunc_{tag}"]),
        "stitch": (O.stitch_adrt(s), g[f"stitch_{tag}"]),
        "stitchrr": (O.stitch_adrt(s, True), g[f"stitchrr_{tag}"]),
    }
    for i in range(O.num_iters(n)):
        checks[f"adrtstep{i}"] = (O.adrt_step(s, i), g[f"adrtstep{i}_{tag}"])
        checks[f"bdrtstep{i}"] = (O.bdrt_step(s, i), g[f"bdrtstep{i}_{tag}"])
    if n >= 2:
        checks["restr"] = (O.press_fmg_restriction(s), g[f"restr_{tag}"])
        checks["highpass"] = (O.press_fmg_highpass(x), g[f"highpass_{tag}"])
        checks["interp"] = (O.interp_to_cart(s), g[f"interp_{tag}"])
        checks["fmgstep"] = (O.iadrt_fmg_step(s), g[f"fmgstep_{tag}"])
    for name, (got, want) in checks.items():
        assert bytes_equal(got, want), f"{name} {tag}: {first_diff(got, want)}"


@pytest.mark.parametrize("dn", list(DTYPES))
@pytest.mark.parametrize("n,B", [(32, 3), (64, 3), (128, 2), (256, 1)])
def test_oracle_vs_golden_hashes(golden_hashes, dn, n, B):
    dt = DTYPES[dn]
    h = golden_hashes[f"{dn}_n{n}_b{B}"]
    x = make_image(1000 + n, (B, n, n), dt)
    y = O.adrt(x)
    s = make_sino(2000 + n, y.shape, dt)
    assert sha(y) == h["adrt"]
    assert sha(O.bdrt(y)) == h["bdrt_of_adrt"]
    assert sha(O.bdrt(s)) == h["bdrt"]
    assert sha(O.iadrt(s)) == h["iadrt"]
    for i in (0, O.num_iters(n) // 2, O.num_iters(n) - 1):
        assert sha(O.adrt_step(s, i)) == h["adrt_step"][i]
        assert sha(O.bdrt_step(s, i)) == h["bdrt_step"][i]
    assert sha(O.interp_to_cart(s)) == h["interp"]
    assert sha(O.press_fmg_restriction(s)) == h["restr"]
    assert sha(O.press_fmg_highpass(x)) == h["highpass"]
    assert sha(O.press_fmg_prolongation(x)) == h["prol"]
    if n <= 64:
        assert sha(O.iadrt_fmg_step(s)) == h["fmgstep"]


def test_oracle_config0(golden_hashes):
    # BASELINE.json configs[0]: adrt.adrt on one 256x256 float32 uniform image
    x = np.random.default_rng(0).random((256, 256), dtype=np.float32)
    assert sha(O.adrt(x)) == golden_hashes["config0_adrt_256_f32_uniform_seed0"]


def test_survey_anchors_oracle():
    # SURVEY.md section 8c anchors (sha256 prefixes of reference outputs)
    import hashlib

    rng = np.random.default_rng(1234)
    x32 = rng.standard_normal((3, 64, 64)).astype(np.float32)
    y = O.adrt(x32)
    assert hashlib.sha256(y.tobytes()).hexdigest()[:16] == "13db36b5fcd5e442"
    assert hashlib.sha256(O.bdrt(y).tobytes()).hexdigest()[:16] == "d3098db331797af5"
    x64 = rng.standard_normal((3, 64, 64))
    y = O.adrt(x64)
    assert hashlib.sha256(y.tobytes()).hexdigest()[:16] == "8c78c3a83075a395"
    assert hashlib.sha256(O.bdrt(y).tobytes()).hexdigest()[:16] == "9e895ce182f326e7"


@pytest.mark.skipif(not ref_loader.have_ref_cdefs(), reason="oracle/_ref not built")
@pytest.mark.parametrize("dn", list(DTYPES))
def test_oracle_vs_live_reference(dn):
    ref = ref_loader.load_ref_cdefs()
    dt = DTYPES[dn]
    for n, B in ((1, 1), (2, 2), (8, 3), (32, 2), (128, 1)):
        x = make_image(7 + n, (B, n, n), dt)
        y = ref.adrt(x)
        s = make_sino(9 + n, y.shape, dt)
        assert bytes_equal(O.adrt(x), y)
        assert bytes_equal(O.bdrt(s), ref.bdrt(s))
        assert bytes_equal(O.iadrt(s), ref.iadrt(s))
        for i in range(O.num_iters(n)):
            assert bytes_equal(O.adrt_step(s, i), ref.adrt_step(s, i))
            assert bytes_equal(O.bdrt_step(s, i), ref.bdrt_step(s, i))
        if n >= 2:
            assert bytes_equal(O.interp_to_cart(s), ref.interp_to_cart(s))
            assert bytes_equal(O.press_fmg_restriction(s), ref.press_fmg_restriction(s))
            assert bytes_equal(O.press_fmg_highpass(x), ref.press_fmg_highpass(x))
        assert bytes_equal(O.press_fmg_prolongation(x), ref.press_fmg_prolongation(x))


def test_helpers_ref_fmg_step_is_the_reference_pass():
    """tests/helpers.ref_fmg_step (the reference's multigrid pass driven by its compiled core,
    used by the config-4 GPU test) is bytes-equal to the reference package's own
    core.iadrt_fmg_step wherever the package can be imported (this container)."""
    from helpers import ref_fmg_step

    if not ref_loader.have_ref_package():
        pytest.skip("reference package not mounted")
    pkg = ref_loader.load_ref_package()
    ref = ref_loader.load_ref_cdefs()
    for dt in (np.float32, np.float64):
        for shape in ((16, 16), (2, 32, 32)):
            a = pkg.adrt(make_image(5, shape, dt))
            want = pkg.core.iadrt_fmg_step(a)
            got = ref_fmg_step(ref, a)
            assert got.tobytes() == want.tobytes()
