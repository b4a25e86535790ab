#!/usr/bin/env python
"""Headline benchmark: ADRT forward + bdrt throughput (Gpixel/s) on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
                    [--batch B] [--n SIDE] [--dtype f32|f64]

One "step" = ``y = adrt(x)`` followed by ``z = bdrt(y)`` on one batch of
synthetic images (BASELINE.json metric).  Default workload is
the headline configuration: a GLOBAL batch of ``64 x 2048^2 float32``; with N > 1
(one process per GPU under torchrun) the batch is sharded ``64 / N`` images per GPU
 and no data-path collective is involved (batch items never
interact), so the run is strong-scaled: value = 64 * n^2 / max-over-ranks time.
``--scaling weak`` keeps ``--batch`` images on every GPU instead; the default run
also measures that configuration for a few steps and reports it as ``weak_scaled``.

Printed JSON (rank 0, one line):
  value        device-resident Gpixel/s (x, y, z live in HBM; CUDA-event timed)
  e2e          same metric through the NumPy-facing host API with pinned host
               buffers (and the out= extension), H2D + D2H inside the timed region;
               e2e.copy_ceiling = the same bytes moved by bare cudaMemcpyAsync on all
               ranks at once (what the host/PCIe side allows), e2e.frac_of_ceiling
  e2e_default  the reference's own calling convention: pageable ndarray in, fresh
               pageable ndarray out (bounce buffers + page faults inside the call)
  roofline     algorithmic bytes (25 n^2 - 12 n) * s * B of one step / measured
               step time, against MEASURED_PEAKS.json's copy bandwidth
  cpu_baseline the unmodified reference (oracle/_ref, C++/OpenMP) on this box's
               host cores, bounded sample of the same workload
``--impl reference`` times only that CPU reference and prints the same line
shape with "impl": "reference".
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "adrt_fwd_plus_bdrt_throughput"
UNIT = "Gpixel/s"


_REAL_STDOUT = None


def quiet_stdout():
    """stdout carries exactly ONE JSON line.  Libraries print there too (NCCL's "NCCL version ..."
    banner comes from C code), so file descriptor 1 points at stderr while the benchmark runs and
    emit() writes the line to the real stdout."""
    global _REAL_STDOUT
    if _REAL_STDOUT is None:
        sys.stdout.flush()
        _REAL_STDOUT = os.dup(1)
        os.dup2(2, 1)


def emit(line: dict):
    sys.stdout.flush()
    data = (json.dumps(line) + "\n").encode()
    if _REAL_STDOUT is None:
        os.write(1, data)
    else:
        os.write(_REAL_STDOUT, data)


def host_cores() -> int:
    try:
        return len(os.sched_getaffinity(0))
    except AttributeError:
        return os.cpu_count() or 1


def algorithmic_bytes(B: int, n: int, itemsize: int) -> int:
    """Compulsory traffic of adrt + bdrt: every input read once, every output
    written once = (n^2 + S) + 2 S elements per image, S = 4 (2n-1) n."""
    return (25 * n * n - 12 * n) * itemsize * B


def measured_peak_gbs():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        with open(p) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


# --------------------------------------------------------------------------------
# CPU reference timing (oracle/_ref = the reference's own C++/OpenMP core)
# --------------------------------------------------------------------------------
def time_reference(n: int, np_dtype, budget_s: float, max_images: int, seed: int = 0):
    """Time adrt_ref.adrt + adrt_ref.bdrt on a bounded batch; returns
    (gpixel_per_s, images, seconds, kind, threads)."""
    import numpy as np

    threads = host_cores()
    # torchrun exports OMP_NUM_THREADS=1 to its workers; the reference arm is meant to
    # use every host core, so set it explicitly (before libgomp initialises).
    if "ADRT_B200_REF_THREADS" in os.environ:
        threads = int(os.environ["ADRT_B200_REF_THREADS"])
    os.environ["OMP_NUM_THREADS"] = str(threads)
    from oracle import ref_loader

    if ref_loader.have_ref_cdefs():
        ref = ref_loader.load_ref_cdefs()
        kind = "reference"
        fwd, bwd = ref.adrt, ref.bdrt
    else:  # the plain-C port (single threaded)
        from oracle import oracle as O

        kind, threads = "port", 1
        fwd, bwd = O.adrt, O.bdrt
    rng = np.random.default_rng(seed)

    def run(b):
        x = rng.random((b, n, n), dtype=np.float32).astype(np_dtype, copy=False)
        t0 = time.perf_counter()
        y = fwd(x)
        z = bwd(y)
        dt = time.perf_counter() - t0
        del y, z
        return dt

    t1 = run(1)  # also warms the thread pool / page faults
    t1 = min(t1, run(1))
    images = int(max(1, min(max_images, budget_s / max(t1, 1e-6))))
    # keep host memory bounded: the reference needs ~3 sinograms per call
    chunk = max(1, min(images, 8))
    total, done = 0.0, 0
    while done < images:
        b = min(chunk, images - done)
        total += run(b)
        done += b
    return images * n * n / total / 1e9, images, total, kind, threads


def shard_sizes(global_batch: int, world: int, scaling: str):
    """Images per rank.  strong: the global batch split as evenly as possible
    (64 / G per GPU); weak: `global_batch` on every rank."""
    if scaling == "weak":
        return [global_batch] * world
    base, rem = divmod(global_batch, world)
    return [base + (1 if r < rem else 0) for r in range(world)]


def reference_arm(args, np_dtype):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    steps = max(1, args.steps)
    total_images = sum(shard_sizes(args.batch, args.gpus, args.scaling))
    per_step_budget = max(2.0, min(20.0, 90.0 / (steps + args.warmup)))
    vals, imgs, secs = [], 0, 0.0
    kind, threads = "reference", host_cores()
    for i in range(args.warmup + steps):
        v, im, s, kind, threads = time_reference(args.n, np_dtype, per_step_budget, total_images, seed=i)
        if i >= args.warmup:
            vals.append(v)
            imgs, secs = im, s
    value = sum(vals) / len(vals)
    ms = 1e3 * (total_images * args.n * args.n / 1e9) / value
    sample = (f"SAMPLED: {imgs} of {total_images} images of {args.n}^2 {args.dtype} timed per step ({secs:.2f} s) and "
              f"extrapolated (per-image cost is constant: batch items are independent, adrt_cdefs_adrt.hpp:71); "
              f"adrt+bdrt, OMP threads={threads}")
    line = {
        "impl": "reference",
        "metric": METRIC, "value": value, "unit": UNIT,
        "n_gpus": args.gpus, "steps": steps, "warmup": args.warmup,
        "ms_per_step": ms, "higher_is_better": True, "scaling": args.scaling,
        "vs_baseline": None, "dtype": args.dtype, "data": "synthetic",
        "config": workload_config(args),
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": kind, "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    emit(line)


def workload_config(args):
    sizes = shard_sizes(args.batch, args.gpus, args.scaling)
    return {
        "workload": f"adrt.adrt + adrt.bdrt on a global batch of {sum(sizes)} x {args.n}^2 {args.dtype}",
        "batch_per_gpu": sizes[0], "n": args.n,
        "global_batch": sum(sizes),
        "parallelism": f"batch-sharded x{args.gpus} ({args.scaling} scaling), no collectives",
        "l2": "inputs larger than L2 (image batch and sinogram are GBs); no explicit flush",
    }


# --------------------------------------------------------------------------------
# clocks sampler
# --------------------------------------------------------------------------------
class ClockSampler:
    REASONS = {
        0x1: "gpu_idle", 0x2: "applications_clocks_setting", 0x4: "sw_power_cap",
        0x8: "hw_slowdown", 0x10: "sync_boost", 0x20: "sw_thermal_slowdown",
        0x40: "hw_thermal_slowdown", 0x80: "hw_power_brake_slowdown", 0x100: "display_clock_setting",
    }

    def __init__(self, index: int):
        self.samples, self.reasons, self.max_mhz = [], set(), None
        self._stop = threading.Event()
        self._thr = None
        try:
            import pynvml

            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self.nv = None

    def _loop(self):
        nv = self.nv
        while not self._stop.is_set():
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                mask = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                for bit, name in self.REASONS.items():
                    if mask & bit and name != "gpu_idle":
                        self.reasons.add(name)
            except Exception:
                pass
            self._stop.wait(0.05)

    def start(self):
        if self.nv is not None:
            self._thr = threading.Thread(target=self._loop, daemon=True)
            self._thr.start()

    def stop(self):
        self._stop.set()
        if self._thr:
            self._thr.join()
        s = sorted(self.samples)
        return {
            "sm_mhz": s[len(s) // 2] if s else None,
            "sm_max_mhz": self.max_mhz,
            "reasons": sorted(self.reasons),
            "samples": len(s),
        }


# --------------------------------------------------------------------------------
# our arm
# --------------------------------------------------------------------------------
def latest_traffic(n: int, dtype: str):
    """DRAM bytes per image (fwd + bdrt) from the newest committed `ncu --set full` capture
    (profiles/rNN_traffic.json, written by tools/ncu_traffic.py together with the commit it was
    taken at).  Only valid for the configuration and kernel set it was captured on."""
    import glob

    for path in sorted(glob.glob(os.path.join(ROOT, "profiles", "r*_traffic.json")), reverse=True):
        try:
            with open(path) as f:
                tj = json.load(f)
            if tj["n"] == n and tj["dtype"] == dtype:
                return tj["dram_bytes_per_image_fwd_plus_bdrt"], {
                    "file": os.path.relpath(path, ROOT), "commit": tj.get("commit"), "batch": tj.get("B")}
        except Exception:
            continue
    return None, None


def ours(args, np_dtype):
    import numpy as np
    import torch

    import adrt_b200 as adrt
    from adrt_b200 import _lib

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if world != args.gpus:
        if world == 1 and args.gpus > 1:
            raise SystemExit("--gpus N > 1 must be launched with torch.distributed.run (one rank per GPU)")
        args.gpus = world
    torch.cuda.set_device(local_rank)
    adrt.set_device(local_rank)
    # one process per GPU: sit on the GPU's NUMA node before any host buffer exists
    from adrt_b200._shard import bind_host_to_device
    numa_bound = bind_host_to_device(local_rank) if world > 1 else False
    dev = torch.device("cuda", local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist

        dist.init_process_group("nccl", device_id=dev)
    lib = _lib.load()
    n = args.n
    sizes = shard_sizes(args.batch, world, args.scaling)
    B, total_images = sizes[rank], sum(sizes)
    tdtype = torch.float32 if args.dtype == "f32" else torch.float64
    itemsize = 4 if args.dtype == "f32" else 8

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(v: float) -> float:
        t = torch.tensor([v], device=dev, dtype=torch.float64)
        if dist is not None:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def device_run(b: int, steps: int, warmup: int, sample_clocks: bool):
        """`steps` timed steps on `b` device-resident images; returns max-over-ranks ms/step etc."""
        g = torch.Generator(device=dev).manual_seed(1234 + rank)
        x = torch.rand((max(b, 1), n, n), device=dev, dtype=tdtype, generator=g)[:b]
        y = torch.empty((b, 4, 2 * n - 1, n), device=dev, dtype=tdtype)
        z = torch.empty_like(y)

        def step():
            if b:
                adrt.adrt(x, out=y)
                adrt.bdrt(y, out=z)

        for _ in range(warmup):
            step()
        barrier()
        # per-transform split (not part of the timed region)
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
        ev[0].record()
        if b:
            adrt.adrt(x, out=y)
        ev[1].record()
        if b:
            adrt.bdrt(y, out=z)
        ev[2].record()
        torch.cuda.synchronize()
        split = (ev[0].elapsed_time(ev[1]), ev[1].elapsed_time(ev[2]))
        sampler = ClockSampler(local_rank) if sample_clocks else None
        barrier()
        if sampler:
            sampler.start()
        launches0 = lib.adrt_b200_launch_count()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            step()
        e1.record()
        barrier()
        launches = lib.adrt_b200_launch_count() - launches0
        clocks = sampler.stop() if sampler else None
        ms_step = max_over_ranks(e0.elapsed_time(e1)) / steps
        return {"ms_step": ms_step, "launches": int(launches), "clocks": clocks, "split": split, "x": x, "y": y}

    warm = max(3, args.warmup)
    main_run = device_run(B, args.steps, warm, True)
    ms_step = main_run["ms_step"]
    value = total_images * n * n / (ms_step * 1e-3) / 1e9
    x, y = main_run.pop("x"), main_run.pop("y")
    t_adrt_ms, t_bdrt_ms = main_run["split"]

    # ---- end to end through the NumPy-facing API ---------------------------------
    e2e, e2e_default = None, None
    if not args.no_e2e and B > 0:
        torch.cuda.empty_cache()
        e2e_steps = max(1, min(args.steps, args.e2e_steps))
        sino_shape = (B, 4, 2 * n - 1, n)
        try:
            hx = torch.empty((B, n, n), dtype=tdtype).pin_memory()
            hy = torch.empty(sino_shape, dtype=tdtype).pin_memory()
            hz = torch.empty(sino_shape, dtype=tdtype).pin_memory()
            hx.copy_(x.cpu())
            nx, ny, nz = hx.numpy(), hy.numpy(), hz.numpy()
            adrt.adrt(nx, out=ny); adrt.bdrt(ny, out=nz)  # warm the staging buffers
            barrier()
            t0 = time.perf_counter()
            for _ in range(e2e_steps):
                adrt.adrt(nx, out=ny)
                adrt.bdrt(ny, out=nz)
            torch.cuda.synchronize()
            dt_step = max_over_ranks(time.perf_counter() - t0) / e2e_steps
            h2d, d2h = int(nx.nbytes + ny.nbytes), int(ny.nbytes + nz.nbytes)
            # what the host side allows: the same bytes, same directions, bare async copies between
            # the same pinned buffers and device memory, all ranks at once (no kernels)
            z = torch.empty_like(y)
            cs = torch.cuda.Stream(device=dev)
            barrier()
            t0 = time.perf_counter()
            for _ in range(e2e_steps):
                # adrt's copies then bdrt's, as the two calls order them: D2H of y, then H2D of y beside D2H of z
                x.copy_(hx, non_blocking=True); hy.copy_(y, non_blocking=True)
                cs.wait_stream(torch.cuda.current_stream())
                with torch.cuda.stream(cs):
                    y.copy_(hy, non_blocking=True)
                hz.copy_(z, non_blocking=True)
                torch.cuda.current_stream().wait_stream(cs)
            torch.cuda.synchronize()
            ceil_step = max_over_ranks(time.perf_counter() - t0) / e2e_steps
            del z
            e2e = {
                "value": total_images * n * n / dt_step / 1e9, "unit": UNIT,
                "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                "ms_per_step": dt_step * 1e3, "steps": e2e_steps,
                "api": "adrt_b200.adrt / adrt_b200.bdrt on pinned numpy.ndarray with out= (adrt_b200_host_adrt / _bdrt)",
                "numa_bound": bool(numa_bound),
                "copy_ceiling": {
                    "value": total_images * n * n / ceil_step / 1e9, "unit": UNIT, "ms_per_step": ceil_step * 1e3,
                    "what": "the same H2D and D2H bytes, in the order the two calls need them (image in, sinogram out; then sinogram in beside sinogram out, full duplex), as bare cudaMemcpyAsync on every rank at once, no kernels",
                    "gbs_per_gpu": (h2d + d2h) / ceil_step / 1e9,
                },
                "frac_of_ceiling": ceil_step / dt_step,
            }
            # end-to-end result check on one image against the device-resident result
            assert np.array_equal(ny[0], y[0].cpu().numpy()), "host path and device path disagree"
            del hx, hy, hz, nz
            # ---- the reference's calling convention: pageable in, fresh pageable out ----
            px = np.array(nx, copy=True)  # ordinary (pageable) ndarray
            del nx, ny
            # one untimed step: the engine hands out large results from a pool of page-locked blocks
            # (adrt_b200/_pinned.py), which the first call of a size has to allocate
            pz = adrt.bdrt(adrt.adrt(px))
            del pz
            barrier()
            t0 = time.perf_counter()
            for _ in range(e2e_steps):
                py = adrt.adrt(px)
                pz = adrt.bdrt(py)
                del pz
            dt_def = max_over_ranks(time.perf_counter() - t0) / e2e_steps
            e2e_default = {
                "value": total_images * n * n / dt_def / 1e9, "unit": UNIT, "ms_per_step": dt_def * 1e3,
                "steps": e2e_steps, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                "api": "y = adrt_b200.adrt(x); z = adrt_b200.bdrt(y) on pageable numpy.ndarray, fresh result arrays "
                       "(the reference's own contract: py.cpp:177 PyArray_SimpleNew); results >= 32 MiB come from the "
                       "engine's pool of page-locked blocks (warm: one untimed step before)",
                "pinned_result_pool": os.environ.get("ADRT_B200_PINNED_RESULTS", "1") != "0",
            }
            assert np.array_equal(py[0], y[0].cpu().numpy()), "default host path and device path disagree"
            del py, px
        except Exception as exc:  # report, never fake
            if e2e is None:
                e2e = {"value": None, "unit": UNIT, "error": repr(exc)[:200]}
            else:
                e2e_default = {"value": None, "unit": UNIT, "error": repr(exc)[:200]}
    del x, y
    torch.cuda.empty_cache()

    # ---- the other scaling flavour, a few steps (extra key; not the headline) ------
    other = None
    if world > 1 and not args.no_other_scaling:
        other_kind = "weak" if args.scaling == "strong" else "strong"
        osz = shard_sizes(args.batch, world, other_kind)
        r = device_run(osz[rank], max(3, min(args.steps, 5)), 3, False)
        other = {"scaling": other_kind, "batch_per_gpu": osz[0], "global_batch": sum(osz),
                 "ms_per_step": r["ms_step"], "value": sum(osz) * n * n / (r["ms_step"] * 1e-3) / 1e9, "unit": UNIT}
        del r

    if rank != 0:
        if dist is not None:
            dist.destroy_process_group()
        return

    peak, peak_src = measured_peak_gbs()
    # roofline of the dominant "kernel" = the whole step on ONE GPU (rank 0's shard)
    abytes = algorithmic_bytes(B, n, itemsize)
    achieved = abytes / (ms_step * 1e-3) / 1e9
    traffic, traffic_src = None, None
    per_image, src = latest_traffic(n, args.dtype)
    if per_image is not None and lib.adrt_b200_get_mode() == 0:
        traffic, traffic_src = per_image * B, src
    roofline = {
        "bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s",
        "frac": achieved / peak, "traffic": traffic, "traffic_source": traffic_src, "peak_source": peak_src,
        "kernel": "whole step (adrt then bdrt kernels of this GPU's shard of the batch)",
        "algorithmic_bytes_per_step": abytes,
        "split": {
            "adrt_ms": t_adrt_ms, "bdrt_ms": t_bdrt_ms,
            "adrt_frac": (n * n + 4 * (2 * n - 1) * n) * itemsize * B / (t_adrt_ms * 1e-3) / 1e9 / peak,
            "bdrt_frac": 2 * 4 * (2 * n - 1) * n * itemsize * B / (t_bdrt_ms * 1e-3) / 1e9 / peak,
        },
    }
    cpu = None
    if world == 1 and not args.no_cpu:
        v, imgs, secs, kind, threads = time_reference(n, np_dtype, args.cpu_budget, B)
        cpu = {
            "value": v, "unit": UNIT, "cores": threads, "kind": kind,
            "sample": f"SAMPLED: {imgs} of {B} images of {n}^2 {args.dtype} ({secs:.2f} s), adrt+bdrt, OMP threads={threads}",
        }
    line = {
        "metric": METRIC, "value": value, "unit": UNIT,
        "n_gpus": world, "steps": args.steps, "warmup": warm,
        "ms_per_step": ms_step, "higher_is_better": True, "scaling": args.scaling,
        "vs_baseline": None, "dtype": args.dtype, "data": "synthetic",
        "config": workload_config(args),
        "clocks": main_run["clocks"], "e2e": e2e, "e2e_default": e2e_default, "gpu_launches": main_run["launches"],
        "roofline": roofline, "cpu_baseline": cpu,
        "mode": "fused" if lib.adrt_b200_get_mode() == 0 else "per-stage",
    }
    if other is not None:
        line[other["scaling"] + "_scaled"] = other
    emit(line)
    if dist is not None:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", choices=["ours", "reference"], default="ours")
    ap.add_argument("--batch", type=int, default=64,
                    help="GLOBAL batch (strong scaling, default) or per-GPU batch (--scaling weak)")
    ap.add_argument("--scaling", choices=["strong", "weak"], default="strong")
    ap.add_argument("--no-other-scaling", action="store_true", help="skip the short run of the other scaling flavour")
    ap.add_argument("--n", type=int, default=2048)
    ap.add_argument("--dtype", choices=["f32", "f64"], default="f32")
    ap.add_argument("--mode", type=int, default=0, help="0 fused (default), 1 per-stage kernels")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--e2e-steps", type=int, default=3)
    ap.add_argument("--cpu-budget", type=float, default=12.0, help="seconds of CPU reference work")
    args = ap.parse_args()
    import numpy as np

    np_dtype = np.float32 if args.dtype == "f32" else np.float64
    quiet_stdout()
    if args.impl == "reference":
        reference_arm(args, np_dtype)
        return
    from adrt_b200 import _lib

    _lib.load().adrt_b200_set_mode(args.mode)
    ours(args, np_dtype)


if __name__ == "__main__":
    main()
