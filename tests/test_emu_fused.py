"""CPU: the fused-pass tile algorithms (adrt_b200/csrc/fused_tile.h, the code the
GPU runs) executed by the host emulator tests/emu/emu_fused.cpp, compared
bit-for-bit with the oracle.  Covers every pass kind (image / workspace / public
layout on either side), radix-4 and radix-2 steps, every stages-per-pass value,
multi-pass splits, masked boundary tiles and signed zeros."""
import ctypes
import os
import subprocess

import numpy as np
import pytest

from helpers import bytes_equal, first_diff, make_image, make_sino
from oracle import oracle as O

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "emu", "emu_fused.cpp")
SO = os.path.join(HERE, "emu", "_build", "libemu.so")


@pytest.fixture(scope="module")
def emu():
    deps = [SRC] + [os.path.join(HERE, "..", "adrt_b200", "csrc", f) for f in ("fused_tile.h", "fused_plan.h", "stream_tile.h", "stage_tile.h", "iadrt_tile.h")]
    # ADRT_EMU_CXXFLAGS="-DMACRO=1 ...": emulate an A/B build of the tile code (tools/build_variant.sh) instead
    extra = os.environ.get("ADRT_EMU_CXXFLAGS", "").split()
    so = SO.replace("libemu.so", "libemu_variant.so") if extra else SO
    if extra or not os.path.exists(so) or any(os.path.getmtime(d) > os.path.getmtime(so) for d in deps):
        os.makedirs(os.path.dirname(so), exist_ok=True)
        subprocess.run(["/usr/bin/g++", "-std=c++17", "-O2", "-ffp-contract=off", "-fPIC", "-shared",
                        "-Wno-unknown-pragmas"] + extra + ["-o", so, SRC], check=True)
    return ctypes.CDLL(so)


def _run(emu, name, a, out_shape):
    suffix = "f32" if a.dtype == np.float32 else "f64"
    a = np.ascontiguousarray(a)
    out = np.full(out_shape, np.nan, dtype=a.dtype)
    B = a.shape[0]
    n = a.shape[-1]
    rc = getattr(emu, f"{name}_{suffix}")(ctypes.c_void_p(a.ctypes.data), ctypes.c_void_p(out.ctypes.data),
                                           ctypes.c_int64(B), ctypes.c_int64(n))
    assert rc == 0
    return out


def _check(emu, n, B, dt, split=None):
    for k in ("ADRT_B200_SPLIT", "ADRT_B200_SPLIT_BDRT"):
        os.environ.pop(k, None)
    if split:
        os.environ["ADRT_B200_SPLIT"] = split
        os.environ["ADRT_B200_SPLIT_BDRT"] = split
    try:
        x = make_image(11 + n, (B, n, n), dt)
        y = _run(emu, "emu_adrt", x, (B, 4, 2 * n - 1, n))
        want = O.adrt(x)
        assert bytes_equal(y, want), f"adrt n={n} split={split}: {first_diff(y, want)}"
        s = make_sino(13 + n, want.shape, dt)
        z = _run(emu, "emu_bdrt", s, s.shape)
        wz = O.bdrt(s)
        assert bytes_equal(z, wz), f"bdrt n={n} split={split}: {first_diff(z, wz)}"
        # all negative zeros: every copy-vs-add decision is visible in the sign bits
        s0 = np.full_like(s, -0.0)
        z = _run(emu, "emu_bdrt", s0, s.shape)
        wz = O.bdrt(s0)
        assert bytes_equal(z, wz), f"bdrt(-0) n={n} split={split}: {first_diff(z, wz)}"
        x0 = np.full_like(x, -0.0)
        y = _run(emu, "emu_adrt", x0, want.shape)
        assert bytes_equal(y, O.adrt(x0)), f"adrt(-0) n={n} split={split}"
    finally:
        for k in ("ADRT_B200_SPLIT", "ADRT_B200_SPLIT_BDRT"):
            os.environ.pop(k, None)


@pytest.mark.parametrize("dt", [np.float32, np.float64])
@pytest.mark.parametrize("n", [2, 4, 8, 16, 32, 64])
def test_single_pass(emu, n, dt):
    _check(emu, n, 2, dt)


@pytest.mark.parametrize("dt", [np.float32, np.float64])
@pytest.mark.parametrize("n", [128, 256])
def test_default_split(emu, n, dt):
    _check(emu, n, 1, dt)


@pytest.mark.parametrize("n,split", [
    (4, "1,1"), (8, "1,2"), (8, "2,1"), (16, "2,2"), (16, "1,3"), (16, "3,1"), (32, "2,3"), (32, "3,2"),
    (32, "1,1,3"), (64, "3,3"), (64, "2,2,2"), (64, "5,1"), (64, "1,5"), (128, "4,3"), (128, "3,4"),
    (128, "2,5"), (128, "5,2"), (128, "3,2,2"), (256, "5,3"), (256, "2,3,3"), (256, "2,2,2,2"),
])
def test_forced_splits(emu, n, split):
    _check(emu, n, 1, np.float32, split)


@pytest.mark.parametrize("dt", [np.float32, np.float64])
def test_six_stage_pass(emu, dt):
    _check(emu, 128, 1, dt, "6,1")
    _check(emu, 128, 1, dt, "1,6")


def test_skewed_rows_near_tile_boundary(emu):
    # regression: a group whose support ends within 3 offsets below a tile boundary of a
    # workspace-storing pass (rows are stored with a skew of up to 3 elements)
    _check(emu, 512, 1, np.float32, "6,2,1")
    _check(emu, 1024, 1, np.float32, "6,3,1")


def test_medium_default(emu):
    # K = 9: two passes (5, 4); several d-tiles per group, masked boundary tiles
    _check(emu, 512, 1, np.float32)


def test_medium_default_f64(emu):
    # fp64 at a size whose image pass has INTERIOR tiles (every window inside the image): the loader
    # that fetches the jobs of two rounds before using either (fwd_radix4_from_image_interior), the
    # paired chunks of the sinogram loader and the unbranched direct first step
    _check(emu, 512, 1, np.float64)
    _check(emu, 512, 1, np.float64, "3,6")
    # the interior loader at the other two group sizes (default plans reach them from 4096^2 up only)
    _check(emu, 512, 1, np.float64, "4,5")
    _check(emu, 512, 1, np.float64, "6,3")


@pytest.mark.parametrize("n,rows,split", [
    (64, 64, None), (128, 128, None), (256, 256, None), (256, 256, "3,5"), (256, 100, "5,3"),
    (512, 512, None), (512, 512, "3,3,3"), (512, 1, None), (1024, 1024, None),
])
def test_bdrt_rows(emu, n, rows, split):
    # row-limited back-projection (adrt_b200_bdrt_rows): offsets d < rows must be
    # bit-identical to the full transform although whole tiles of every pass are skipped
    os.environ.pop("ADRT_B200_SPLIT_BDRT", None)
    if split:
        os.environ["ADRT_B200_SPLIT_BDRT"] = split
    try:
        s = make_sino(17 + n, (1, 4, 2 * n - 1, n), np.float32)
        out = np.full(s.shape, np.nan, dtype=s.dtype)
        rc = emu.emu_bdrt_rows_f32(ctypes.c_void_p(s.ctypes.data), ctypes.c_void_p(out.ctypes.data),
                                   ctypes.c_int64(1), ctypes.c_int64(n), ctypes.c_int64(rows))
        assert rc == 0
        want = O.bdrt(s)
        assert bytes_equal(out[:, :, :rows], want[:, :, :rows]), first_diff(out[:, :, :rows], want[:, :, :rows])
        if rows <= n // 2:
            assert np.isnan(out[:, :, -1]).all()   # the far rows were really skipped
    finally:
        os.environ.pop("ADRT_B200_SPLIT_BDRT", None)


@pytest.mark.parametrize("dt", [np.float32, np.float64])
@pytest.mark.parametrize("n,rows,split", [(128, 128, None), (256, 256, None), (256, 511, "4,4"), (512, 512, None), (1024, 1024, None)])
def test_bdrt_subtract_on_load(emu, n, rows, split, dt):
    # iadrt_fmg_step's residual (core.py:329): bdrt(a - b) with the subtraction done by the loader of the
    # first pass (BwdProgram<..., kSub>) == bdrt of the separately subtracted sinogram, bit for bit,
    # masked boundary tiles and signed zeros included (a == b gives +0.0 where a copy would keep -0.0)
    os.environ.pop("ADRT_B200_SPLIT_BDRT", None)
    if split:
        os.environ["ADRT_B200_SPLIT_BDRT"] = split
    try:
        suffix = "f32" if dt == np.float32 else "f64"
        a = make_sino(41 + n, (1, 4, 2 * n - 1, n), dt)
        b = make_sino(43 + n, (1, 4, 2 * n - 1, n), dt)
        b.reshape(-1)[::7] = a.reshape(-1)[::7]
        a.reshape(-1)[::11] = -0.0
        out = np.full(a.shape, np.nan, dtype=dt)
        rc = getattr(emu, f"emu_bdrt_sub_{suffix}")(ctypes.c_void_p(a.ctypes.data), ctypes.c_void_p(b.ctypes.data),
                                                    ctypes.c_void_p(out.ctypes.data), ctypes.c_int64(1), ctypes.c_int64(n),
                                                    ctypes.c_int64(rows))
        if rc == 4:
            pytest.skip("the plan of this size has no public-layout first pass with a fused step")
        assert rc == 0
        want = O.bdrt(a - b)
        assert bytes_equal(out[:, :, :rows], want[:, :, :rows]), first_diff(out[:, :, :rows], want[:, :, :rows])
    finally:
        os.environ.pop("ADRT_B200_SPLIT_BDRT", None)


# ---------------------------------------------------------------------------------------------
# Streaming passes (adrt_b200/csrc/stream_tile.h: fp32, 5 or 6 stages, butterflies with register
# history, in-place tiles).  ADRT_B200_STREAM_SET=all forces every pass kind through them; each case
# runs twice, with the threads of a phase emulated in ascending and in descending order -- a result
# that depended on the order would mean two threads race on a tile cell inside a phase.
def _check_stream(emu, n, split, rows=None):
    emu.emu_stream_tiles.restype = ctypes.c_longlong
    keys = ("ADRT_B200_SPLIT", "ADRT_B200_SPLIT_BDRT", "ADRT_B200_STREAM_SET", "ADRT_B200_STAGE_SET")
    for k in keys:
        os.environ.pop(k, None)
    os.environ["ADRT_B200_STREAM_SET"] = "all"
    os.environ["ADRT_B200_STAGE_SET"] = ""      # the staged variants (stage_tile.h) have their own tests below
    if split:
        os.environ["ADRT_B200_SPLIT"] = split
        os.environ["ADRT_B200_SPLIT_BDRT"] = split
    try:
        x = make_image(11 + n, (1, n, n), np.float32)
        want = O.adrt(x)
        s = make_sino(13 + n, want.shape, np.float32)
        wz = O.bdrt(s)
        x0, s0 = np.full_like(x, -0.0), np.full_like(s, -0.0)
        wy0, wz0 = O.adrt(x0), O.bdrt(s0)
        for order in (0, 1):
            emu.emu_set_order(order)
            t0 = emu.emu_stream_tiles()
            y = _run(emu, "emu_adrt", x, want.shape)
            assert emu.emu_stream_tiles() > t0, "no streaming tile ran"
            assert bytes_equal(y, want), f"adrt n={n} split={split} order={order}: {first_diff(y, want)}"
            assert bytes_equal(_run(emu, "emu_adrt", x0, want.shape), wy0), f"adrt(-0) n={n} split={split} order={order}"
            t0 = emu.emu_stream_tiles()
            z = _run(emu, "emu_bdrt", s, s.shape)
            assert emu.emu_stream_tiles() > t0, "no streaming tile ran"
            assert bytes_equal(z, wz), f"bdrt n={n} split={split} order={order}: {first_diff(z, wz)}"
            assert bytes_equal(_run(emu, "emu_bdrt", s0, s.shape), wz0), f"bdrt(-0) n={n} split={split} order={order}"
            if rows:
                out = np.full(s.shape, np.nan, dtype=s.dtype)
                rc = emu.emu_bdrt_rows_f32(ctypes.c_void_p(s.ctypes.data), ctypes.c_void_p(out.ctypes.data),
                                           ctypes.c_int64(1), ctypes.c_int64(n), ctypes.c_int64(rows))
                assert rc == 0
                assert bytes_equal(out[:, :, :rows], wz[:, :, :rows]), first_diff(out[:, :, :rows], wz[:, :, :rows])
    finally:
        emu.emu_set_order(0)
        for k in keys:
            os.environ.pop(k, None)


@pytest.mark.parametrize("n,split,rows", [
    (32, None, None), (64, None, 20),                 # one pass: image / sinogram on both sides, every tile masked
    (128, "5,2", None), (128, "2,5", 128),            # 5 stages next to the image / next to the public layout
    (256, "6,2", None), (256, "2,6", 100),            # 6 stages likewise
    (512, "6,3", 512), (512, "3,6", None),            # several d-tiles per group: interior and boundary tiles
    (1024, None, 1024), (1024, "5,5", None),          # default plan (5 + 5), both orders of the split
])
def test_streaming_passes(emu, n, split, rows):
    _check_stream(emu, n, split, rows)


@pytest.mark.parametrize("n,split", [(128, "1,6"), (256, "2,6"), (512, "3,6")])
def test_streaming_two_threads_per_segment(emu, n, split, monkeypatch):
    # ADRT_B200_STREAM_SPLIT2: the six-stage forward pass that stores the public layout on 128 threads per tile
    # (every segment shared by two threads); both thread orders expose a race on the in-place tile
    monkeypatch.setenv("ADRT_B200_STREAM_SPLIT2", "1")
    _check_stream(emu, n, split)


def test_streaming_two_six_stage_passes(emu):
    # K = 12 = 6 + 6 (the 4096^2 plan) at the smallest size that has it is too slow to emulate;
    # 6 + 5 and 5 + 6 (the 2048^2 plans) are covered through 11-stage splits of n = 2048 on the GPU.
    _check_stream(emu, 512, "6,3")
    _check_stream(emu, 512, "4,5")


# ---------------------------------------------------------------------------------------------
# Staged passes (adrt_b200/csrc/stage_tile.h): the five-stage fp32 passes next to the images / the
# public-layout sinogram with their input tile delivered by tensor-map or bulk copies into a staging
# buffer and an out-of-place first butterfly step.  ADRT_B200_STAGE_SET selects them; both thread orders.
def _check_staged(emu, n, split, rows=None, B=1):
    emu.emu_staged_tiles.restype = ctypes.c_longlong
    keys = ("ADRT_B200_SPLIT", "ADRT_B200_SPLIT_BDRT", "ADRT_B200_STAGE_SET")
    for k in keys:
        os.environ.pop(k, None)
    os.environ["ADRT_B200_STAGE_SET"] = "f5p,b5p"
    if split:
        # forward order in both: the staged pass is the first forward pass and undoes the last one
        os.environ["ADRT_B200_SPLIT"] = split
        os.environ["ADRT_B200_SPLIT_BDRT"] = ",".join(reversed(split.split(",")))
    try:
        x = make_image(11 + n, (B, n, n), np.float32)
        want = O.adrt(x)
        s = make_sino(13 + n, want.shape, np.float32)
        wz = O.bdrt(s)
        x0, s0 = np.full_like(x, -0.0), np.full_like(s, -0.0)
        wy0, wz0 = O.adrt(x0), O.bdrt(s0)
        for order in (0, 1):
            emu.emu_set_order(order)
            t0 = emu.emu_staged_tiles()
            y = _run(emu, "emu_adrt", x, want.shape)
            assert emu.emu_staged_tiles() > t0, "no staged tile ran"
            assert bytes_equal(y, want), f"adrt n={n} split={split} order={order}: {first_diff(y, want)}"
            assert bytes_equal(_run(emu, "emu_adrt", x0, want.shape), wy0), f"adrt(-0) n={n} split={split} order={order}"
            t0 = emu.emu_staged_tiles()
            z = _run(emu, "emu_bdrt", s, s.shape)
            assert emu.emu_staged_tiles() > t0, "no staged tile ran"
            assert bytes_equal(z, wz), f"bdrt n={n} split={split} order={order}: {first_diff(z, wz)}"
            assert bytes_equal(_run(emu, "emu_bdrt", s0, s.shape), wz0), f"bdrt(-0) n={n} split={split} order={order}"
            if rows:
                out = np.full(s.shape, np.nan, dtype=s.dtype)
                rc = emu.emu_bdrt_rows_f32(ctypes.c_void_p(s.ctypes.data), ctypes.c_void_p(out.ctypes.data),
                                           ctypes.c_int64(B), ctypes.c_int64(n), ctypes.c_int64(rows))
                assert rc == 0
                assert bytes_equal(out[:, :, :rows], wz[:, :, :rows]), first_diff(out[:, :, :rows], wz[:, :, :rows])
    finally:
        emu.emu_set_order(0)
        for k in keys:
            os.environ.pop(k, None)


@pytest.mark.parametrize("n,split,rows,B", [
    (64, "5,1", None, 2),                  # one group, every tile masked, two images (plane -> image / quadrant)
    (128, "5,2", 100, 1),                  # the staged pass first in both directions
    (256, "5,3", None, 1),
    (512, "5,4", 512, 1),                  # several d-tiles per group: interior and boundary tiles, row-limited
    (1024, None, None, 1),                 # the 1024^2 plan (5 + 5)
])
def test_staged_passes(emu, n, split, rows, B):
    _check_staged(emu, n, split, rows, B)


# ---------------------------------------------------------------------------------------------
# Fused normal operator: adrt hands its result to bdrt as R-layout rows
# (fused_plan.h rows_out / rows_in) -- the public-layout store of the last forward pass and the
# public-layout load of the first transposed pass never happen.
def _check_normal(emu, n, dt, rows, stream_set=None, split=None):
    keys = ("ADRT_B200_SPLIT", "ADRT_B200_SPLIT_BDRT", "ADRT_B200_STREAM_SET")
    for k in keys:
        os.environ.pop(k, None)
    if stream_set is not None:
        os.environ["ADRT_B200_STREAM_SET"] = stream_set
    if split:
        os.environ["ADRT_B200_SPLIT"] = split
        os.environ["ADRT_B200_SPLIT_BDRT"] = split
    try:
        suffix = "f32" if dt == np.float32 else "f64"
        D = 2 * n - 1
        pitch = (D + 3) & ~3
        for x in (make_image(23 + n, (1, n, n), dt), np.full((1, n, n), -0.0, dtype=dt)):
            mid = np.full((1, 4, n, pitch), np.nan, dtype=dt)
            out = np.full((1, 4, D, n), np.nan, dtype=dt)
            rc = getattr(emu, f"emu_normal_{suffix}")(ctypes.c_void_p(x.ctypes.data), ctypes.c_void_p(mid.ctypes.data),
                                                       ctypes.c_void_p(out.ctypes.data), ctypes.c_int64(1), ctypes.c_int64(n),
                                                       ctypes.c_int64(rows))
            assert rc == 0
            y = O.adrt(x)
            # the hand-off buffer is the sinogram transposed (row = angle), complete up to offset D
            want_mid = np.ascontiguousarray(np.swapaxes(y, -1, -2))
            assert bytes_equal(mid[..., :D], want_mid), f"rows-layout adrt n={n}: {first_diff(mid[..., :D], want_mid)}"
            want = O.bdrt(y)
            assert bytes_equal(out[:, :, :rows], want[:, :, :rows]), f"normal n={n} rows={rows}: {first_diff(out[:, :, :rows], want[:, :, :rows])}"
    finally:
        for k in keys:
            os.environ.pop(k, None)


@pytest.mark.parametrize("dt", [np.float32, np.float64])
@pytest.mark.parametrize("n", [2, 8, 32, 64, 128, 256])
def test_normal_operator_rows_handoff(emu, n, dt):
    _check_normal(emu, n, dt, n)
    if n <= 64:
        _check_normal(emu, n, dt, 2 * n - 1)


@pytest.mark.parametrize("n,split,stream_set", [
    (256, "3,5", "all"), (256, "5,3", "all"), (512, "3,6", "all"), (512, "6,3", "all"), (1024, None, None), (512, "3,3,3", ""),
])
def test_normal_operator_rows_handoff_streaming(emu, n, split, stream_set):
    _check_normal(emu, n, np.float32, n, stream_set=stream_set, split=split)


# ---------------------------------------------------------------------------------------------
# Angle-block sharding of one image over several ranks (fused_plan.h part_*): every rank runs
# the passes before the last one on its own image-row blocks, ONE exchange of workspace rows, then the last
# pass on its own angle range (bdrt: the mirror image).  The emulator plays the ranks one after the other
# with NaN-filled private buffers, so a row the protocol forgets to send shows up as NaN.
@pytest.mark.parametrize("dt", [np.float32, np.float64])
@pytest.mark.parametrize("n,parts,m_last", [(16, 2, 1), (32, 2, 3), (64, 4, 2), (128, 2, 3), (256, 2, 1), (256, 4, 3), (512, 8, 3)])
def test_angle_block_sharding(emu, n, parts, m_last, dt):
    suffix = "f32" if dt == np.float32 else "f64"
    x = make_image(41 + n, (1, n, n), dt)
    y = np.full((1, 4, 2 * n - 1, n), np.nan, dtype=dt)
    rc = getattr(emu, f"emu_adrt_parts_{suffix}")(ctypes.c_void_p(x.ctypes.data), ctypes.c_void_p(y.ctypes.data),
                                                   ctypes.c_int64(1), ctypes.c_int64(n), ctypes.c_int(parts), ctypes.c_int(m_last))
    assert rc == 0
    want = O.adrt(x)
    assert bytes_equal(y, want), f"sharded adrt n={n} parts={parts}: {first_diff(y, want)}"
    s = make_sino(43 + n, want.shape, dt)
    for rows in (2 * n - 1, n):
        z = np.full(s.shape, np.nan, dtype=dt)
        rc = getattr(emu, f"emu_bdrt_parts_{suffix}")(ctypes.c_void_p(s.ctypes.data), ctypes.c_void_p(z.ctypes.data),
                                                       ctypes.c_int64(1), ctypes.c_int64(n), ctypes.c_int(parts), ctypes.c_int(m_last),
                                                       ctypes.c_int64(rows))
        assert rc == 0
        wz = O.bdrt(s)
        assert bytes_equal(z[:, :, :rows], wz[:, :, :rows]), f"sharded bdrt n={n} parts={parts} rows={rows}: {first_diff(z[:, :, :rows], wz[:, :, :rows])}"


# ---------------------------------------------------------------------------------------------
# Fused multi-stage iadrt passes (adrt_b200/csrc/iadrt_tile.h): warp-level sweeps in pre-shifted frames,
# levels one row apart, rings in shared memory, column-major workspace between passes.  The emulator
# plays the 32 lanes of every warp phase by phase (ascending and descending lane order).
@pytest.mark.parametrize("dt", [np.float32, np.float64])
@pytest.mark.parametrize("n,split", [
    (2, None), (4, None), (8, None), (16, None), (32, None),          # one pass of 1..5 stages
    (8, "1,2"), (16, "2,2"), (32, "3,2"), (32, "1,1,3"), (64, "3,3"),  # workspace hand-off, 2 and 3 passes
    (64, None), (128, None), (256, "4,4"), (256, "5,3"), (512, None), (1024, "5,5"),
])
def test_fused_iadrt(emu, n, split, dt):
    os.environ.pop("ADRT_B200_IADRT_SPLIT", None)
    if split:
        os.environ["ADRT_B200_IADRT_SPLIT"] = split
    try:
        B = 2 if n <= 64 else 1
        for s in (make_sino(29 + n, (B, 4, 2 * n - 1, n), dt), np.full((B, 4, 2 * n - 1, n), -0.0, dtype=dt)):
            want = O.iadrt(s)
            for order in (0, 1):
                emu.emu_set_order(order)
                got = _run(emu, "emu_iadrt", s, s.shape)
                assert bytes_equal(got, want), f"iadrt n={n} split={split} order={order}: {first_diff(got, want)}"
    finally:
        emu.emu_set_order(0)
        os.environ.pop("ADRT_B200_IADRT_SPLIT", None)


def test_bdrt_zero_tile_skipping(emu):
    """ADRT_B200_SKIP_ZERO=1 (opt-in): the producer of a streaming transposed pass does not write its all-zero
    tiles and the streaming loader clips every workspace row at its support D - a*j; the emulator's
    workspaces start as NaN, so any read of an unwritten cell would surface."""
    os.environ["ADRT_B200_SKIP_ZERO"] = "1"
    try:
        _check_stream(emu, 512, "3,6", 300)
        _check_stream(emu, 1024, None, 1024)
        _check(emu, 256, 1, np.float32)
    finally:
        os.environ.pop("ADRT_B200_SKIP_ZERO", None)
