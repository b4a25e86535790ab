"""GPU: parity of the CUDA path (through the C ABI) with the oracle, the
committed golden vectors and -- when oracle/_ref is present -- the reference's
own compiled core.  Bit-exact (bytes-equal) for everything but the iterated
multigrid inverse."""
import numpy as np
import pytest

import adrt_b200 as adrt
from adrt_b200 import _adrt_cdefs as cd
from adrt_b200 import _lib
from helpers import DTYPES, bytes_equal, first_diff, make_image, make_sino, sha
from oracle import oracle as O
from oracle import ref_loader

pytestmark = pytest.mark.gpu

SMALL = [(dn, n, B) for dn in DTYPES for n in (1, 2, 4, 8, 16) for B in (0, 2)]


def _eq(got, want, what):
    assert bytes_equal(got, want), f"{what}: {first_diff(got, want)}"


@pytest.fixture(params=[0, 1], ids=["fused", "per-stage"])
def mode(request):
    lib = _lib.load()
    lib.adrt_b200_set_mode(request.param)
    yield request.param
    lib.adrt_b200_set_mode(0)


@pytest.mark.parametrize("dn,n,B", SMALL)
def test_small_golden_all_ops(golden_small, mode, dn, n, B):
    g = golden_small
    tag = f"{dn}_n{n}_b{B}"
    x, s = g[f"x_{tag}"], g[f"s_{tag}"]
    _eq(adrt.adrt(x), g[f"adrt_{tag}"], "adrt")
    _eq(adrt.bdrt(s), g[f"bdrt_{tag}"], "bdrt")
    _eq(adrt.iadrt(s), g[f"iadrt_{tag}"], "iadrt")
    _eq(cd.adrt_init(x), g[f"init_{tag}"], "adrt_init")
    for i in range(O.num_iters(n)):
        _eq(adrt.core.adrt_step(s, i), g[f"adrtstep{i}_{tag}"], f"adrt_step {i}")
        _eq(adrt.core.bdrt_step(s, step=i), g[f"bdrtstep{i}_{tag}"], f"bdrt_step {i}")
    _eq(cd.press_fmg_prolongation(x), g[f"prol_{tag}"], "prolongation")
    if n >= 2:
        _eq(cd.press_fmg_restriction(s), g[f"restr_{tag}"], "restriction")
        _eq(cd.press_fmg_highpass(x), g[f"highpass_{tag}"], "highpass")
        _eq(adrt.utils.interp_to_cart(s), g[f"interp_{tag}"], "interp_to_cart")
        _eq(adrt.core.iadrt_fmg_step(s), g[f"fmgstep_{tag}"], "iadrt_fmg_step")


@pytest.mark.parametrize("dn", list(DTYPES))
@pytest.mark.parametrize("n,B", [(32, 3), (64, 3), (128, 2), (256, 1), (512, 1)])
def test_golden_hashes(golden_hashes, mode, dn, n, B):
    dt = DTYPES[dn]
    h = golden_hashes[f"{dn}_n{n}_b{B}"]
    x = make_image(1000 + n, (B, n, n), dt)
    y = adrt.adrt(x)
    s = make_sino(2000 + n, y.shape, dt)
    assert sha(y) == h["adrt"], first_diff(y, O.adrt(x))
    z = adrt.bdrt(y)
    assert sha(z) == h["bdrt_of_adrt"], first_diff(z, O.bdrt(y))
    z = adrt.bdrt(s)
    assert sha(z) == h["bdrt"], first_diff(z, O.bdrt(s))
    if mode == 0:
        assert sha(adrt.iadrt(s)) == h["iadrt"]
        for i in range(O.num_iters(n)):
            assert sha(adrt.core.adrt_step(s, i)) == h["adrt_step"][i], f"adrt_step {i}"
            assert sha(adrt.core.bdrt_step(s, i)) == h["bdrt_step"][i], f"bdrt_step {i}"
        assert sha(adrt.utils.interp_to_cart(s)) == h["interp"]
        assert sha(cd.press_fmg_restriction(s)) == h["restr"]
        assert sha(cd.press_fmg_highpass(x)) == h["highpass"]
        assert sha(cd.press_fmg_prolongation(x)) == h["prol"]
        assert sha(adrt.core.iadrt_fmg_step(s)) == h["fmgstep"]


def test_config0_bit_exact(golden_hashes):
    # BASELINE.json configs[0]
    x = np.random.default_rng(0).random((256, 256), dtype=np.float32)
    assert sha(adrt.adrt(x)) == golden_hashes["config0_adrt_256_f32_uniform_seed0"]


def test_survey_anchors():
    import hashlib

    rng = np.random.default_rng(1234)
    x32 = rng.standard_normal((3, 64, 64)).astype(np.float32)
    y = adrt.adrt(x32)
    assert hashlib.sha256(y.tobytes()).hexdigest()[:16] == "13db36b5fcd5e442"
    assert hashlib.sha256(adrt.bdrt(y).tobytes()).hexdigest()[:16] == "d3098db331797af5"
    x64 = rng.standard_normal((3, 64, 64))
    y = adrt.adrt(x64)
    assert hashlib.sha256(y.tobytes()).hexdigest()[:16] == "8c78c3a83075a395"
    assert hashlib.sha256(adrt.bdrt(y).tobytes()).hexdigest()[:16] == "9e895ce182f326e7"


@pytest.mark.parametrize("dn", list(DTYPES))
@pytest.mark.parametrize("n,B", [(1024, 2), (2048, 1), (4096, 1), (8192, 1)])
def test_large_vs_oracle(mode, dn, n, B):
    """Sizes of BASELINE configs 1-5 on a batch the CPU oracle finishes in seconds
    (8192 = three fused passes)."""
    if n == 8192 and mode == 1:
        pytest.skip("per-stage path is covered up to 4096")
    dt = DTYPES[dn]
    x = make_image(31 + n, (B, n, n), dt)
    if ref_loader.have_ref_cdefs():
        ref = ref_loader.load_ref_cdefs()
        want_y = ref.adrt(x)
        y = adrt.adrt(x)
        _eq(y, want_y, f"adrt n={n}")
        _eq(adrt.bdrt(y), ref.bdrt(want_y), f"bdrt n={n}")
    else:
        y = adrt.adrt(x)
        _eq(y, O.adrt(x), f"adrt n={n}")
        _eq(adrt.bdrt(y), O.bdrt(y), f"bdrt n={n}")


def test_many_planes(mode):
    """More (image, quadrant) planes than gridDim.z allows: kernels loop over planes."""
    for n in (4, 16):
        B = 17000  # 68000 planes > 65535
        x = make_image(8, (B, n, n), np.float32)
        y = adrt.adrt(x)
        _eq(y, O.adrt(x), f"adrt B={B} n={n}")
        _eq(adrt.bdrt(y), O.bdrt(y), f"bdrt B={B} n={n}")
        if mode == 0:
            _eq(adrt.core.adrt_step(y, 1), O.adrt_step(y, 1), "adrt_step many planes")
            _eq(adrt.iadrt(y[:100]), O.iadrt(y[:100]), "iadrt")


def test_special_values(mode):
    """NaN / Inf propagate, negative zeros keep their sign (SURVEY 8a exactness)."""
    n = 32
    x = np.full((n, n), -0.0, dtype=np.float32)
    _eq(adrt.adrt(x), O.adrt(x), "all -0.0 adrt")
    s = np.full((4, 2 * n - 1, n), -0.0, dtype=np.float32)
    _eq(adrt.bdrt(s), O.bdrt(s), "all -0.0 bdrt")
    x = make_image(5, (n, n), np.float64)
    x[3, 7] = np.nan
    x[9, 1] = np.inf
    x[20, 30] = -np.inf
    got, want = adrt.adrt(x), O.adrt(x)
    assert np.array_equal(np.isnan(got), np.isnan(want))
    assert np.array_equal(got[~np.isnan(got)], want[~np.isnan(want)])
    sub = np.full((n, n), 1e-42, dtype=np.float32)  # subnormals are not flushed
    _eq(adrt.adrt(sub), O.adrt(sub), "subnormal adrt")


def test_torch_path_matches_numpy_path(mode):
    import torch

    for dt in (np.float32, np.float64):
        x = make_image(77, (3, 64, 64), dt)
        xt = torch.from_numpy(x).cuda()
        yt = adrt.adrt(xt)
        assert isinstance(yt, torch.Tensor) and yt.is_cuda and tuple(yt.shape) == (3, 4, 127, 64)
        y = adrt.adrt(x)
        _eq(yt.cpu().numpy(), y, "torch adrt")
        _eq(adrt.bdrt(yt).cpu().numpy(), adrt.bdrt(y), "torch bdrt")
        _eq(adrt.iadrt(yt).cpu().numpy(), adrt.iadrt(y), "torch iadrt")
        _eq(adrt.utils.interp_to_cart(yt).cpu().numpy(), adrt.utils.interp_to_cart(y), "torch interp")
        _eq(adrt.core.adrt_init(xt).cpu().numpy(), adrt.core.adrt_init(x), "torch init")
        _eq(adrt.utils.truncate(yt).cpu().numpy(), adrt.utils.truncate(y), "torch truncate")
        _eq(cd.truncate(yt).cpu().numpy(), adrt.utils.truncate(y), "kernel truncate")
        _eq(adrt.utils.stitch_adrt(yt).cpu().numpy(), adrt.utils.stitch_adrt(y), "torch stitch")
        # single image (no batch dim) and DLPack producer other than torch.Tensor
        _eq(adrt.adrt(xt[0]).cpu().numpy(), y[0], "torch adrt unbatched")


def test_iter_generators(mode):
    x = make_image(3, (2, 16, 16), np.float32)
    items = list(adrt.core.adrt_iter(x))
    assert len(items) == 5
    _eq(items[0], adrt.core.adrt_init(x), "iter[0]")
    _eq(items[-1], adrt.adrt(x), "last(adrt_iter) == adrt")
    assert all(not it.flags.writeable for it in adrt.core.adrt_iter(x, copy=False))
    y = items[-1]
    bits = list(adrt.core.bdrt_iter(y))
    assert len(bits) == 4
    np.testing.assert_array_equal(bits[-1], adrt.bdrt(y))


def test_adjoint_small(mode):
    """A == (truncate . bdrt)^T by materialising both (reference tests/test_bdrt.py:310-325)."""
    for n in (1, 2, 4, 8):
        D = 2 * n - 1
        eye_img = np.eye(n * n, dtype=np.float64).reshape(n * n, n, n)
        A = adrt.adrt(eye_img).reshape(n * n, -1)                  # rows: pixels
        eye_sino = np.eye(4 * D * n, dtype=np.float64).reshape(4 * D * n, 4, D, n)
        Bm = adrt.utils.truncate(adrt.bdrt(eye_sino))              # (4Dn, 4, n, n)
        # quadrant q of the back-projection only sees quadrant q of the input
        Bt = np.zeros((n * n, 4 * D * n))
        for q in range(4):
            blk = slice(q * D * n, (q + 1) * D * n)
            Bt[:, blk] = Bm[blk, q].reshape(D * n, n * n).T
        np.testing.assert_array_equal(A, Bt)


def test_fmg_inverse_tolerance():
    """iadrt_fmg iterates vs the oracle, rel. tol 1e-5 (fp32) / 1e-12 (fp64)
    as stated in BASELINE.json's north_star."""
    for dt, tol in ((np.float32, 1e-5), (np.float64, 1e-12)):
        n = 64
        yy, xx = np.mgrid[0:n, 0:n]
        img = np.exp(-((xx - 40.0) ** 2 + (yy - 24.0) ** 2) / 60.0).astype(dt)
        a = O.adrt(img)
        want = O.iadrt_fmg_iter(a, 3)
        it = adrt.core.iadrt_fmg_iter(a)
        for k in range(3):
            got = next(it)
            err = np.linalg.norm(got - want[k]) / np.linalg.norm(want[k])
            assert err <= tol, (dt, k, err)
        res = adrt.iadrt_fmg(a, max_iters=3)
        assert res.shape == (n, n) and res.flags.writeable
        assert np.linalg.norm(res - img) / np.linalg.norm(img) < 0.2


def test_fmg_step_batch_bit_exact():
    """BASELINE config 4 shape family: batched iadrt_fmg_step on device equals the
    oracle's restatement of core.py:318-331 bit for bit (every level's adrt/bdrt,
    restriction, prolongation, high-pass and the NumPy-ordered quadrant mean)."""
    import torch

    for dt in (np.float32, np.float64):
        n, B = 256, 2
        a = O.adrt(make_image(91, (B, n, n), dt))
        want = O.iadrt_fmg_step(a)
        _eq(adrt.core.iadrt_fmg_step(a), want, "fmg_step numpy path")
        got_t = adrt.core.iadrt_fmg_step(torch.from_numpy(a).cuda())
        _eq(got_t.cpu().numpy(), want, "fmg_step tensor path")


def test_cg_recipe_matches_scipy_on_oracle():
    """docs/examples.cginverse.md:40-67: CG on the normal equations; compare with
    SciPy's cg driven by the CPU oracle operators (tolerance: both stop at rtol 1e-6)."""
    from scipy.sparse.linalg import LinearOperator, cg

    from adrt_b200 import recipes

    n = 16
    xs = np.linspace(-1, 1, n)
    xx, yy = np.meshgrid(xs, xs)
    img = np.exp(-6 * (xx ** 2 + (yy - 0.2) ** 2)).astype(np.float64)
    b = O.adrt(img)

    def matvec(v):
        y = O.bdrt(O.adrt(v.reshape(n, n)))
        return O.truncate(y).mean(axis=0).ravel()

    A = LinearOperator((n * n, n * n), matvec=matvec, dtype=np.float64)
    want, info = cg(A, O.truncate(O.bdrt(b)).mean(axis=0).ravel(), rtol=1e-10, atol=0.0)
    assert info == 0
    got, iters = recipes.iadrt_cg(b, rtol=1e-10, return_info=True)
    assert iters > 3
    np.testing.assert_allclose(got, want.reshape(n, n), rtol=1e-6, atol=1e-8)
    np.testing.assert_allclose(got, img, rtol=1e-5, atol=1e-6)
    # the operator itself is bit-exact apart from NumPy's pairwise mean vs our sequential one
    import torch

    v = make_image(4, (n, n), np.float64)
    nv = recipes.normal_operator(torch.from_numpy(v).cuda()).cpu().numpy()
    np.testing.assert_allclose(nv, matvec(v.ravel()).reshape(n, n), rtol=1e-13, atol=1e-13)
    with pytest.raises(ValueError, match="batch dimension not supported"):
        recipes.iadrt_cg(np.zeros((2, 4, 31, 16)))


def test_quadrant_subsets_and_sharded_operator():
    """Single-image sharding building blocks: any quadrant range of adrt equals the
    slice of the full transform, bdrt of a plane subset equals the slice of bdrt, and
    the quadrant-sharded normal operator (world of 1 here) equals the fused one."""
    import torch

    from adrt_b200 import _shard, recipes

    for dt in (np.float32, np.float64):
        x = make_image(21, (2, 128, 128), dt)
        xt = torch.from_numpy(x).cuda()
        full = adrt.adrt(xt)
        back = adrt.bdrt(full)
        for qf, qc in ((0, 4), (0, 1), (1, 1), (2, 2), (3, 1), (1, 3)):
            sub = cd.adrt_quadrants(xt, qf, qc)
            assert torch.equal(sub.view(torch.uint8), full[:, qf:qf + qc].contiguous().view(torch.uint8)), (qf, qc)
            bsub = cd.bdrt_planes(sub)
            assert torch.equal(bsub.view(torch.uint8), back[:, qf:qf + qc].contiguous().view(torch.uint8)), (qf, qc)
        one = cd.adrt_quadrants(xt[0], 2, 1)
        assert tuple(one.shape) == (1, 255, 128)
        got = _shard.sharded_normal_operator(xt[0])
        want = recipes.normal_operator(xt[0])
        assert torch.equal(got.view(torch.uint8), want.view(torch.uint8))
    with pytest.raises(ValueError, match="bad quadrant range"):
        cd.adrt_quadrants(xt, 3, 2)


@pytest.mark.parametrize("dn", list(DTYPES))
@pytest.mark.parametrize("n,B", [(2, 3), (16, 2), (64, 2), (256, 2), (1024, 1), (2048, 2), (8192, 1)])
def test_bdrt_rows_and_truncate_mean(dn, n, B):
    """adrt_b200_bdrt_rows: offsets d < rows are bit-identical to the full bdrt, and the
    fused normal-operator tail equals truncate_mean(bdrt(.)) (utils.py:231-242,
    core.py:329)."""
    import torch

    dt = DTYPES[dn]
    s = torch.from_numpy(make_sino(31 + n, (B, 4, 2 * n - 1, n), dt)).cuda()
    full = adrt.bdrt(s)
    for rows in sorted({1, n // 2 + 1, n, 2 * n - 1}):
        part = cd.bdrt_planes(s, rows=rows)
        assert torch.equal(part[:, :, :rows].contiguous().view(torch.uint8),
                           full[:, :, :rows].contiguous().view(torch.uint8)), rows
    got = cd.bdrt_truncate_mean(s, n - 1 if n > 1 else 1)
    want = cd.truncate_mean(full, n - 1 if n > 1 else 1)
    assert torch.equal(got.view(torch.uint8), want.view(torch.uint8))
    if n <= 1024:
        ref = np.mean(O.truncate(O.bdrt(s.cpu().numpy())) / dt(n - 1 if n > 1 else 1), axis=-3)
        _eq(got.cpu().numpy(), ref.astype(dt), "bdrt_truncate_mean vs oracle")
    with pytest.raises(_lib.ADRTB200Error, match="rows"):
        cd.bdrt_planes(s, rows=2 * n)


@pytest.mark.parametrize("dn", list(DTYPES))
def test_fmg_operators_shapes_and_alignment(dn):
    """Press FMG operators on odd / non-square / large shapes (scalar and vector kernel
    variants) and on tensors whose storage is not 16-byte aligned."""
    import torch

    dt = DTYPES[dn]
    for shape in ((3, 5, 7), (2, 6, 10), (1, 33, 64), (2, 2, 2), (1, 1024, 2048), (5, 64, 3), (3, 9, 4), (2, 17, 8), (2, 2, 12)):
        x = make_image(41 + shape[-1], shape, dt)
        _eq(cd.press_fmg_prolongation(x), O.press_fmg_prolongation(x), f"prolongation {shape}")
        if shape[-1] >= 2 and shape[-2] >= 2:
            _eq(cd.press_fmg_highpass(x), O.press_fmg_highpass(x), f"highpass {shape}")
    for n, B in ((2, 3), (4, 2), (8, 1), (64, 2), (1024, 1)):
        s = make_sino(43 + n, (B, 4, 2 * n - 1, n), dt)
        _eq(cd.press_fmg_restriction(s), O.press_fmg_restriction(s), f"restriction n={n}")
    # misaligned storage: views starting one element into a larger buffer
    n = 64
    s = make_sino(47, (2, 4, 2 * n - 1, n), dt)
    x = make_image(48, (2, n, n), dt)
    for arr, fn, ofn in ((s, cd.press_fmg_restriction, O.press_fmg_restriction),
                         (x, cd.press_fmg_prolongation, O.press_fmg_prolongation),
                         (x, cd.press_fmg_highpass, O.press_fmg_highpass),
                         (s, lambda t, **kw: cd.adrt_step(t, 2, **kw), lambda a: O.adrt_step(a, 2)),
                         (s, lambda t, **kw: cd.bdrt_step(t, 2, **kw), lambda a: O.bdrt_step(a, 2))):
        buf = torch.zeros(arr.size + 1, dtype=torch.from_numpy(arr).dtype, device="cuda")
        view = buf[1:].view(arr.shape)
        view.copy_(torch.from_numpy(arr))
        want = ofn(arr)
        obuf = torch.zeros(want.size + 1, dtype=view.dtype, device="cuda")
        oview = obuf[1:].view(want.shape)
        got = fn(view, out=oview)
        _eq(got.cpu().numpy(), want, "misaligned in/out")
        a2 = torch.from_numpy(arr).cuda()
        d = cd.sub(view, a2, out=oview) if want.shape == arr.shape else None
        if d is not None:
            assert not d.any()


@pytest.mark.parametrize("dn", list(DTYPES))
@pytest.mark.parametrize("n,B", [(1, 2), (2, 3), (8, 2), (64, 3), (512, 2), (2048, 1)])
def test_native_fmg_step_equals_composed_operators(mode, dn, n, B):
    """adrt_b200_fmg_step (all levels in one native call) against the same sequence written
    with the public operators as in core.py:318-331."""
    import torch

    dt = DTYPES[dn]
    a = torch.from_numpy(make_sino(51 + n, (B, 4, 2 * n - 1, n), dt)).cuda()
    got = cd.fmg_step(a)
    stack, cur = [], a
    for _ in range(O.num_iters(n)):
        stack.append(cur)
        cur = cd.press_fmg_restriction(cur)
    ret = cur[..., 0, :, :].contiguous()
    m = 1
    while stack:
        m *= 2
        ret = cd.press_fmg_prolongation(ret)
        resid = cd.sub(cd.adrt(ret), stack.pop())
        grad = cd.truncate_mean(cd.bdrt(resid), m - 1)
        ret = cd.sub(ret, cd.press_fmg_highpass(grad))
    assert tuple(got.shape) == (B, n, n)
    assert torch.equal(got.view(torch.uint8), ret.view(torch.uint8))
    if n <= 64:
        _eq(got.cpu().numpy(), O.iadrt_fmg_step(a.cpu().numpy()), "fmg_step vs oracle")


@pytest.mark.parametrize("dn", list(DTYPES))
@pytest.mark.parametrize("n,B", [(128, 3), (1024, 2), (2048, 1)])
def test_step_kernels_large_vs_oracle(dn, n, B):
    """iadrt (wide stages: lane-skewed odd columns through the shared-memory ring, pair loads),
    bdrt_step (2-D CTAs) and adrt_step against the oracle at sizes where those kernels are chosen."""
    dt = DTYPES[dn]
    s = make_sino(77 + n, (B, 4, 2 * n - 1, n), dt)
    _eq(adrt.iadrt(s), O.iadrt(s), f"iadrt n={n}")
    K = O.num_iters(n)
    for i in sorted({0, 1, K // 2, K - 2, K - 1}):
        _eq(adrt.core.bdrt_step(s, i), O.bdrt_step(s, i), f"bdrt_step {i} n={n}")
        _eq(adrt.core.adrt_step(s, i), O.adrt_step(s, i), f"adrt_step {i} n={n}")


def test_step_kernels_misaligned_views():
    """Device tensors that start at an odd element offset take the scalar kernels."""
    import torch

    n = 128
    s = make_sino(5, (2, 4, 2 * n - 1, n), np.float32)
    flat = torch.zeros(s.size + 1, dtype=torch.float32, device="cuda")
    flat[1:] = torch.from_numpy(s).reshape(-1).cuda()
    view = flat[1:].reshape(s.shape)
    _eq(adrt.iadrt(view).cpu().numpy(), O.iadrt(s), "iadrt misaligned")
    _eq(adrt.core.bdrt_step(view, 2).cpu().numpy(), O.bdrt_step(s, 2), "bdrt_step misaligned")
    # the fused transforms use 16/32-byte vector accesses: the shim copies such views
    _eq(adrt.bdrt(view).cpu().numpy(), O.bdrt(s), "bdrt misaligned")
    x = make_image(6, (2, n, n), np.float32)
    flat = torch.zeros(x.size + 3, dtype=torch.float32, device="cuda")
    flat[3:] = torch.from_numpy(x).reshape(-1).cuda()
    _eq(adrt.adrt(flat[3:].reshape(x.shape)).cpu().numpy(), O.adrt(x), "adrt misaligned")
    obuf = torch.zeros(2 * 4 * (2 * n - 1) * n + 1, device="cuda")
    oview = obuf[1:].reshape(2, 4, 2 * n - 1, n)
    got = adrt.adrt(torch.from_numpy(x).cuda(), out=oview)
    assert got.data_ptr() == oview.data_ptr()
    _eq(oview.cpu().numpy(), O.adrt(x), "adrt into a misaligned out")
    # the C ABI itself refuses such pointers for the fused entry points instead of faulting
    lib = _lib.load()
    ws = torch.empty(int(lib.adrt_b200_adrt_workspace_bytes(2, n, 0)), dtype=torch.uint8, device="cuda")
    rc = lib.adrt_b200_adrt(flat[3:].data_ptr(), oview.data_ptr(), 2, n, 0, ws.data_ptr(), ws.numel(), None)
    assert rc != 0 and "aligned" in _lib.last_error()


def test_iadrt_roundtrip():
    # reference tests/test_iadrt.py:185-223
    for n in (16, 32):
        x = np.arange(n * n, dtype=np.float32).reshape(n, n)
        inv = adrt.utils.truncate(adrt.iadrt(adrt.adrt(x)))
        assert np.allclose(inv.mean(axis=0), x)


def test_full_size_properties():
    """BASELINE headline size 64 x 2048^2 fp32 on device: fused == per-stage path
    bit for bit, per-angle mass conservation, and the first images against the
    compiled reference."""
    import torch

    lib = _lib.load()
    B, n = 64, 2048
    g = torch.Generator(device="cuda").manual_seed(0)
    x = torch.randn((B, n, n), device="cuda", dtype=torch.float32, generator=g)
    lib.adrt_b200_set_mode(0)
    y = adrt.adrt(x)
    assert tuple(y.shape) == (B, 4, 2 * n - 1, n)
    z = adrt.bdrt(y)
    lib.adrt_b200_set_mode(1)
    try:
        # per-stage path in chunks of 8 images to bound scratch memory
        for b0 in range(0, B, 8):
            y1 = adrt.adrt(x[b0:b0 + 8])
            assert torch.equal(y1.view(torch.int32), y[b0:b0 + 8].view(torch.int32)), f"adrt chunk {b0}"
            z1 = adrt.bdrt(y1)
            assert torch.equal(z1.view(torch.int32), z[b0:b0 + 8].view(torch.int32)), f"bdrt chunk {b0}"
            del y1, z1
    finally:
        lib.adrt_b200_set_mode(0)
    # every pixel lies on exactly one digital line per angle: column sums == image sum
    xi = torch.randint(0, 8, (2, n, n), device="cuda", generator=g).to(torch.float32)
    yi = adrt.adrt(xi)
    tot = xi.sum(dim=(1, 2), dtype=torch.float64)
    col = yi.sum(dim=2, dtype=torch.float64)  # (2, 4, n)
    assert torch.equal(col, tot[:, None, None].expand_as(col))
    if ref_loader.have_ref_cdefs():
        ref = ref_loader.load_ref_cdefs()
        xs = x[:2].cpu().numpy()
        want = ref.adrt(xs)
        _eq(y[:2].cpu().numpy(), want, "adrt[:2] vs reference")
        _eq(z[:2].cpu().numpy(), ref.bdrt(want), "bdrt[:2] vs reference")


@pytest.mark.parametrize("n,B", [(1024, 2), (2048, 1), (4096, 1)])
def test_streaming_kinds_agree(n, B):
    """Every pass kind through the streaming kernels (stream_tile.h), none of them, and the default
    mix give the same bytes -- signed zeros, NaN and the row-limited back-projection included."""
    import os

    x = make_image(77 + n, (B, n, n), np.float32)
    x[0, 3, 5] = np.nan
    x[0, 7, :4] = np.inf
    want_y = O.adrt(x) if n <= 2048 else None
    s = make_sino(78 + n, (B, 4, 2 * n - 1, n), np.float32)
    want_z = O.bdrt(s) if n <= 2048 else None
    s0 = np.full_like(s, -0.0)
    got = {}
    try:
        for tag, val in (("none", ""), ("all", "all"), ("default", None)):
            os.environ.pop("ADRT_B200_STREAM_SET", None)
            os.environ.pop("ADRT_B200_STAGE_SET", None)
            if val is not None:
                os.environ["ADRT_B200_STREAM_SET"] = val
                os.environ["ADRT_B200_STAGE_SET"] = ""   # the staged variants: test_staged_passes_equal_default
            got[tag] = (adrt.adrt(x), adrt.bdrt(s), adrt.bdrt(s0))
    finally:
        os.environ.pop("ADRT_B200_STREAM_SET", None)
        os.environ.pop("ADRT_B200_STAGE_SET", None)
    for tag in ("all", "default"):
        for i, what in enumerate(("adrt", "bdrt", "bdrt(-0)")):
            _eq(got[tag][i], got["none"][i], f"{what} n={n} streaming set {tag} vs fused_tile.h kernels")
    if want_y is not None:
        # NaN payloads are not compared with the CPU oracle (the GPU's FADD returns the canonical NaN);
        # everything else, signed zeros and infinities included, is bytes-equal
        def canon(a):
            a = a.copy()
            a[np.isnan(a)] = np.float32(np.nan)
            return a

        assert np.isnan(got["all"][0]).any() and np.isinf(got["all"][0]).any()
        _eq(canon(got["all"][0]), canon(want_y), f"adrt n={n} streaming vs oracle")
        _eq(got["all"][1], want_z, f"bdrt n={n} streaming vs oracle")
