#!/usr/bin/env python3
"""Benchmark of the hot path BASELINE.json names: batched factorisations per second.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload chol16] [--impl ours|reference]
                    [--scaling weak|strong] [--headline-only]

Headline workload (the top-level keys of the JSON line): BASELINE.json configs[1] - batched Cholesky fp32
N=16, batch 1M synthetic SPD matrices (genA.m formula) per B200.  One "step" = one pass of the
factorisation over the whole batch.  With N>1 (torchrun, one rank per GPU) every rank factors its own
batch of the same size (weak scaling, plain batch splitting, no collective on the data path; NCCL only
carries the barrier and the per-rank counters); `--scaling strong` splits ONE batch of that size over the
ranks instead.

One JSON line on stdout (rank 0):
  value        factorizations/s, inputs resident in HBM, device-timed (CUDA events), max over ranks
  e2e          same metric through the C ABI's host-pointer entry point (hlsd_*_host) with pinned HOST
               buffers, H2D and D2H copies inside the timed region; `copy_ceiling` is the same staging
               pipeline with the kernel left out (hlsd_selftest_copy_host), measured on all ranks at once
  roofline     algorithmic HBM bytes (or tensor flops) of the dominant kernel / its measured duration, vs
               MEASURED_PEAKS.json
  cpu_baseline the reference's own C++ (oracle/_ref, g++ + HLS header stand-ins) on the host cores (N=1)
  workloads    the other BASELINE.json configurations (LU 64, QR 128, Cholesky and LU 1024), each with the
               same value / roofline / e2e objects (shorter runs)
  per_n        BASELINE.json's metric "per N (4-1024)": one row per (op, N), resident, single calls
`--impl reference` times only the CPU reference (rank 0) and prints the same line shape.
"""
import argparse
import ctypes
import json
import os
import statistics
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "factorizations/sec"

WORKLOADS = {
    # name: (kind, rows, cols, batch per GPU, design)
    "chol16": ("chol", 16, 16, 1 << 20, "cholesky_v1.3"),
    "lu64": ("lu", 64, 64, 262144, "lu_v1.1"),
    # the same elimination with the pivot search switched off (P[k] = k): BASELINE.json words its LU config as
    # "unpivoted"; the reference designs themselves pivot, so lu64 above is the parity workload
    "lu64np": ("lu", 64, 64, 262144, "nopivot"),
    "qr128": ("qr", 128, 128, 16384, "qr_v2.1"),
    "chol32": ("chol", 32, 32, 1 << 19, "cholesky_v1.3"),
    "chol64": ("chol", 64, 64, 262144, "cholesky_v1.3"),
    "chol1024": ("chol", 1024, 1024, 4096, "cholesky_v1.3"),
    "lu1024": ("lu", 1024, 1024, 4096, "lu_v1.1"),
}
HEADLINE = "chol16"
EXTRA_WORKLOADS = ("lu64", "qr128", "chol1024", "lu1024")  # BASELINE.json configs[2..4]
PROF_CLASSES = ("other", "update", "diag", "trsm", "panel", "permute", "exact")


# ------------------------------------------------------------------------------------------------
# multi-GPU plan: plain batch splitting (also exercised on CPU by tests/test_sharding.py)
# ------------------------------------------------------------------------------------------------
def shard_batch(batch, rank, world):
    """Contiguous, balanced [start, end) span of `batch` units for `rank` of `world`."""
    base, extra = divmod(batch, world)
    start = rank * base + min(rank, extra)
    return start, start + base + (1 if rank < extra else 0)


def rank_batch(batch, rank, world, scaling):
    """Units this rank factors: its own full batch (weak) or its span of one shared batch (strong)."""
    if scaling == "weak":
        return batch
    start, end = shard_batch(batch, rank, world)
    return end - start


def combine_ranks(units, ms, device):
    """Whole-job view: units summed over ranks, time = slowest rank."""
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()):
        return units, ms
    t_units = torch.tensor([float(units)], dtype=torch.float64, device=device)
    t_ms = torch.tensor([float(ms)], dtype=torch.float64, device=device)
    dist.all_reduce(t_units, op=dist.ReduceOp.SUM)
    dist.all_reduce(t_ms, op=dist.ReduceOp.MAX)
    return int(round(t_units.item())), t_ms.item()


def algorithmic_bytes(kind, rows, cols):
    """HBM bytes one factorisation must move: operand in + every output array out."""
    if kind == "chol":
        return 2 * rows * rows * 4
    if kind == "lu":
        return 3 * rows * rows * 4 + rows * 4
    return (2 * rows * cols + rows * rows) * 4


def flops(kind, rows, cols):
    if kind == "chol":
        return rows ** 3 / 3.0
    if kind == "lu":
        return 2.0 * rows ** 3 / 3.0
    # Givens: each rotation = 6 flops per rotated pair over the remaining R columns and all of Q's rows
    return sum((rows - 1 - c) * 6.0 * ((cols - 1 - c) + rows) for c in range(cols))


def update_flops_executed(kind, n):
    """Flops the tcgen05 trailing-update kernel executes per matrix (3xTF32: three MMAs per product):
    Cholesky left-looking: tile (r, q) takes q tile products, r >= q; LU in Crout order (lu_ll_kernel): panel q's
    column form (p - q row tiles) and row form (p - 1 - q column tiles) each take q tile products per tile."""
    p = n // 128
    tiles = sum((p - q) * q for q in range(p))
    if kind == "lu":
        tiles += sum((p - 1 - q) * q for q in range(p))
    return 3.0 * tiles * 2.0 * 128 ** 3


# ------------------------------------------------------------------------------------------------
# clocks: NVML poller running during the timed regions
# ------------------------------------------------------------------------------------------------
class ClockSampler:
    REASONS = {0x4: "sw_power_cap", 0x8: "hw_slowdown", 0x20: "sw_thermal_slowdown", 0x40: "hw_thermal_slowdown",
               0x80: "hw_power_brake_slowdown", 0x2: "applications_clocks_setting", 0x10: "sync_boost"}

    def __init__(self, index):
        self.samples, self.reasons, self.max_mhz, self._stop, self._on = [], set(), None, False, False
        self.affinity = "unchanged"
        self.all_cpus = os.sched_getaffinity(0)
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv, self.h = pynvml, pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
            self.thread = threading.Thread(target=self._run, daemon=True)
            self.thread.start()
        except Exception as exc:  # noqa: BLE001
            self.nv, self.err = None, str(exc)

    def bind_cpu_near_gpu(self):
        """Run this process (and the pinned allocations it makes) on the CPUs NVML lists as local to the GPU."""
        if not self.nv:
            return
        try:
            self.nv.nvmlDeviceSetCpuAffinity(self.h)
            self.affinity = f"nvmlDeviceSetCpuAffinity: {len(os.sched_getaffinity(0))} of {len(self.all_cpus)} CPUs"
        except Exception as exc:  # noqa: BLE001
            self.affinity = f"unchanged ({type(exc).__name__})"

    def unbind_cpu(self):
        try:
            os.sched_setaffinity(0, self.all_cpus)
        except Exception:  # noqa: BLE001
            pass

    def _run(self):
        while not self._stop:
            if self._on:
                try:
                    self.samples.append(self.nv.nvmlDeviceGetClockInfo(self.h, self.nv.NVML_CLOCK_SM))
                    mask = self.nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                    for bit, name in self.REASONS.items():
                        if mask & bit:
                            self.reasons.add(name)
                except Exception:  # noqa: BLE001
                    pass
            time.sleep(0.004)

    def region(self, on):
        self._on = on

    def summary(self):
        self._stop = True
        if not self.nv:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvml unavailable"]}
        med = statistics.median(self.samples) if self.samples else None
        return {"sm_mhz": med, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons), "samples": len(self.samples)}


# ------------------------------------------------------------------------------------------------
# CPU reference (oracle/_ref preferred, oracle port otherwise): the only place bench.py runs oracle/
# ------------------------------------------------------------------------------------------------
class CpuReference:
    """One bounded sample of the workload, factored by the reference's own C++ (oracle/_ref) when it
    was built for this size, else by the oracle port (oracle/restate.c), on all host cores."""

    def __init__(self, kind, rows, cols, design, target_s):
        sys.path.insert(0, os.path.join(ROOT, "oracle"))
        import resource
        try:
            resource.setrlimit(resource.RLIMIT_STACK, (resource.RLIM_INFINITY, resource.RLIM_INFINITY))
        except (ValueError, OSError):
            pass
        os.environ.setdefault("OMP_STACKSIZE", "64M")
        import ref
        import restate
        from hls_designs_b200.operand import gen_full_rank, gen_spd
        self.kind, self.rows, self.cols, self.design = kind, rows, cols, design
        self.cores = len(os.sched_getaffinity(0))
        tag = rows if kind != "qr" else f"{rows}x{cols}"
        self.use_ref = ref.available(design, tag)
        self._ref, self._restate = ref, restate
        self._make = (lambda b: gen_spd(rows, batch=b, seed=1)) if kind == "chol" else \
                     (lambda b: gen_full_rank(rows, cols, batch=b, seed=1))
        probe = max(self.cores * 4, 64) if rows < 512 else self.cores
        while True:  # grow the probe until it runs long enough to extrapolate from
            self.A = self._make(probe)
            dt = max(self.run_once(), 1e-6)
            if dt >= 0.25 or probe >= (1 << 20):
                break
            probe *= 4
        count = int(min(max(self.cores, target_s * probe / dt), 1 << 21))
        self.count = max(self.cores, count // self.cores * self.cores)
        self.A = self._make(self.count)

    def run_once(self):
        t0 = time.perf_counter()
        if self.use_ref:
            {"chol": self._ref.ref_cholesky, "lu": self._ref.ref_lu, "qr": self._ref.ref_qr}[self.kind](
                self.A, self.design, threads=self.cores)
        elif self.kind == "chol":
            self._restate.cholesky(self.A, self.design, threads=self.cores)
        elif self.kind == "lu":
            self._restate.lu(self.A, self.design != "nopivot", threads=self.cores)
        else:
            self._restate.qr(self.A, self.design, threads=self.cores)
        return time.perf_counter() - t0

    def describe(self, value, seconds):
        impl = "oracle/_ref (the reference's C++ built with g++ and HLS header stand-ins)" if self.use_ref \
            else "oracle/restate.c (port)"
        return {"value": value, "unit": METRIC, "cores": self.cores, "kind": "reference" if self.use_ref else "port",
                "sample": f"{self.count} matrices of the same workload ({self.rows}x{self.cols} {self.kind}, "
                          f"{self.design}) per pass, {seconds:.1f} s per pass, {impl}, {self.cores} OpenMP threads"}


def cpu_reference_rate(kind, rows, cols, design, target_s):
    cr = CpuReference(kind, rows, cols, design, target_s)
    passes, total = 0, 0.0
    while passes == 0 or (total < target_s and passes < 50):
        total += cr.run_once()
        passes += 1
    out = cr.describe(cr.count * passes / total, total / passes)
    out["sample"] += f", {passes} passes ({total:.1f} s of CPU time in all)"
    return out


def workload_config(name, path, scaling):
    kind, rows, cols, batch, design = WORKLOADS[name]
    per = "per GPU" if scaling == "weak" else "in all, split over the GPUs"
    return {"workload": f"batched {kind} fp32 {rows}x{cols}, batch {batch} {per}, design {design}, synthetic genA operands",
            "batch": batch, "rows": rows, "cols": cols, "design": design, "kernel_path": path,
            "l2": "inputs larger than L2 (no flush needed)"}


# ------------------------------------------------------------------------------------------------
# one workload on this rank's GPU: resident leg, roofline, e2e leg, CPU baseline
# ------------------------------------------------------------------------------------------------
class Runner:
    def __init__(self, args, rank, world, local_rank):
        import torch
        import hls_designs_b200 as hd
        self.torch, self.hd, self.args = torch, hd, args
        self.rank, self.world, self.local_rank = rank, world, local_rank
        self.dev = torch.device("cuda", local_rank)
        self.lib = hd.load()
        self.lib.hlsd_set_device(local_rank)
        self.peaks = {}
        try:
            self.peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:  # noqa: BLE001
            pass
        try:
            self.traffic = json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))
        except Exception:  # noqa: BLE001
            self.traffic = {}

    def barrier(self):
        import torch.distributed as dist
        if self.world > 1:
            dist.barrier()
        self.torch.cuda.synchronize()

    def operand(self, kind, rows, cols, batch, seed):
        torch = self.torch
        g = torch.Generator(device=self.dev).manual_seed(seed)
        if kind == "chol":
            A = torch.rand((batch, rows, rows), generator=g, device=self.dev)
            A = 0.5 * (A + A.transpose(1, 2))
            A.diagonal(dim1=1, dim2=2).add_(float(rows))
            return A.contiguous()
        return torch.randint(1, 101, (batch, rows, cols), generator=g, device=self.dev).float()

    def pinned(self, nbytes, dtype, shape):
        import numpy as np
        p = self.lib.hlsd_host_alloc(nbytes)
        if not p:
            raise MemoryError(self.lib.hlsd_last_error().decode())
        arr = np.ctypeslib.as_array((ctypes.c_byte * nbytes).from_address(p)).view(dtype).reshape(shape)
        return p, arr

    def run(self, name, steps, warmup, e2e_steps, cpu_seconds, sampler):
        import numpy as np
        torch, hd, args = self.torch, self.hd, self.args
        kind, rows, cols, full_batch, design = WORKLOADS[name]
        if args.batch and name == args.workload:
            full_batch = args.batch
        batch = rank_batch(full_batch, self.rank, self.world, args.scaling)
        A = self.operand(kind, rows, cols, batch, 1234 + self.rank)
        fn = {"chol": hd.cholesky, "lu": hd.lu, "qr": hd.qr}[kind]
        fused = bool(args.fused)

        def step():
            return fn(A, design=design, path=args.path, fused=fused)

        for _ in range(max(3, warmup)):
            out = step()
        self.barrier()
        stream = torch.cuda.current_stream()
        evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
        launches0 = hd.launch_count()
        sampler.region(True)
        t_begin, t_end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t_begin.record(stream)
        for s_ev, e_ev in evs:
            s_ev.record(stream)
            out = step()
            e_ev.record(stream)
        t_end.record(stream)
        torch.cuda.synchronize()
        sampler.region(False)
        self.barrier()
        launches = hd.launch_count() - launches0
        total_ms = t_begin.elapsed_time(t_end)
        kern_ms = [s.elapsed_time(e) for s, e in evs]
        units, worst_ms = combine_ranks(batch * steps, total_ms, self.dev)
        value = units / (worst_ms * 1e-3)
        res = {"name": name, "config": workload_config(name, args.path, args.scaling), "value": value, "unit": METRIC,
               "steps": steps, "ms_per_step": worst_ms / steps, "gpu_launches": int(launches),
               "batch_this_rank": batch}
        if fused:
            res["config"]["flags"] = "HLSD_FLAG_FUSED"

        # ---- roofline of the dominant kernel ----
        peaks = self.peaks
        peak_gbs = float(peaks.get("hbm_gbs", 6650.0))
        peak_src = "MEASURED_PEAKS.json hbm_gbs (measured)" if "hbm_gbs" in peaks else "fallback 6650 GB/s (B200_PROFILING.md)"
        med_ms = statistics.median(kern_ms)
        bytes_per_launch = algorithmic_bytes(kind, rows, cols) * batch
        achieved = bytes_per_launch / (med_ms * 1e-3) / 1e9
        gflops = flops(kind, rows, cols) * batch / (med_ms * 1e-3) / 1e9
        tensor = kind in ("chol", "lu") and rows >= 256 and args.path in ("auto", "tensor")
        if not tensor:
            traffic = self.traffic.get(name) if batch == WORKLOADS[name][3] else None
            res["roofline"] = {"bound": "hbm", "achieved": achieved, "peak": peak_gbs, "unit": "GB/s", "frac": achieved / peak_gbs,
                               "traffic": traffic,
                               "traffic_source": "constant from the committed ncu --set full capture (profiles/traffic.json), not "
                                                 "measured in this run" if traffic else None,
                               "peak_source": peak_src, "kernel_ms_median": med_ms, "kernel_ms_min": min(kern_ms),
                               "kernel_ms_max": max(kern_ms), "algorithmic_bytes_per_launch": bytes_per_launch, "gflops": gflops}
        else:
            # blocked tensor-core path: a step is ~25-60 launches.  Per-class kernel times come from the library's own
            # CUDA-event brackets (hlsd_profile_begin/_end) over a few extra steps; the dominant tensor kernel is the
            # trailing update, whose EXECUTED flops (3 MMAs per product, 3xTF32) are set against the tf32 peak
            # (= half the measured bf16 rate).
            peak_tf = float(peaks.get("bf16_tflops_sustained", 1400.0)) / 2.0
            prof_steps = 3
            ms = (ctypes.c_double * len(PROF_CLASSES))()
            cnt = (ctypes.c_int64 * len(PROF_CLASSES))()
            torch.cuda.synchronize()
            self.lib.hlsd_profile_begin()
            for _ in range(prof_steps):
                step()
            torch.cuda.synchronize()
            self.lib.hlsd_profile_end(ms, cnt, len(PROF_CLASSES))
            by_class = {c: ms[i] / prof_steps for i, c in enumerate(PROF_CLASSES) if cnt[i]}
            share = {c: v / sum(by_class.values()) for c, v in by_class.items()}
            upd_ms = by_class.get("update", float("nan"))
            upd_tf = update_flops_executed(kind, rows) * batch / (upd_ms * 1e-3) / 1e12
            res["roofline"] = {"bound": "tensor", "achieved": upd_tf, "peak": peak_tf, "unit": "TFLOP/s", "frac": upd_tf / peak_tf,
                               "traffic": None, "kernel": "blk_ll2_kernel" if kind == "chol" else "lu_ll_kernel<LL_COL|LL_ROW>",
                               "kernel_ms_per_step": upd_ms, "kernel_share_of_step": share.get("update"),
                               "ms_per_step_by_kernel_class": by_class, "share_by_kernel_class": share,
                               "whole_step": {"useful_tflops": gflops / 1e3, "frac_of_tf32_peak": gflops / 1e3 / peak_tf,
                                              "hbm_gbs": achieved, "frac_of_hbm_peak": achieved / peak_gbs},
                               "kernel_ms_median": med_ms, "kernel_ms_min": min(kern_ms), "kernel_ms_max": max(kern_ms),
                               "gflops": gflops,
                               "peak_source": "MEASURED_PEAKS.json bf16_tflops_sustained / 2 (tf32 runs at half the bf16 rate)",
                               "note": "achieved = flops the trailing-update kernel executes per step (3 MMAs per product, 3xTF32) / "
                                       "its own CUDA-event time per step, measured by the library's per-launch event brackets"}

        # ---- end to end through the host-pointer C entry point, pinned host buffers ----
        if not args.no_e2e and e2e_steps > 0:
            res["e2e"] = self.e2e(name, kind, rows, cols, design, batch, A, out, e2e_steps, sampler)

        if res["roofline"]["bound"] == "hbm":
            # context for the exact kernels that are not HBM-bound: the reference's arithmetic has no fused
            # multiply-add (every float operation is rounded separately), so the arithmetic ceiling is one
            # operation per FP32 lane and clock, not two
            props = torch.cuda.get_device_properties(self.dev)
            ghz = (sampler.max_mhz or 1965) / 1e3
            peak = props.multi_processor_count * 128 * ghz / 1e3  # Tflop/s without FMA
            res["roofline"]["fp32_unfused"] = {"achieved": gflops / 1e3, "peak": peak, "unit": "TFLOP/s", "frac": gflops / 1e3 / peak,
                                               "note": "useful flops of the factorisation / (SMs x 128 FP32 lanes x max clock); "
                                                       "the exact kernels issue separately rounded multiplies and adds"}
        del A, out
        torch.cuda.empty_cache()
        if self.rank == 0 and self.world == 1 and not args.no_cpu_baseline and cpu_seconds > 0:
            sampler.unbind_cpu()
            res["cpu_baseline"] = cpu_reference_rate(kind, rows, cols, design, cpu_seconds)
        return res

    def e2e(self, name, kind, rows, cols, design, batch, A, out, e2e_steps, sampler):
        import numpy as np
        torch, hd, lib, args = self.torch, self.hd, self.lib, self.args
        in_mat = rows * cols * 4
        out_mat = {"chol": rows * rows * 4, "lu": 2 * rows * rows * 4 + rows * 4, "qr": (rows * rows + rows * cols) * 4}[kind]
        cap = int(args.e2e_max_gb * (1 << 30))
        e2e_batch = int(max(1, min(batch, cap // (in_mat + out_mat))))
        in_bytes = e2e_batch * in_mat
        out_bytes = e2e_batch * out_mat
        sampler.bind_cpu_near_gpu()  # pinned pages on the NUMA node next to this rank's GPU
        pA, hA = self.pinned(in_bytes, np.float32, (e2e_batch, rows, cols))
        torch.from_numpy(hA).copy_(A[:e2e_batch])
        if kind == "chol":
            shapes = [((e2e_batch, rows, rows), np.float32)]
        elif kind == "lu":
            shapes = [((e2e_batch, rows, rows), np.float32), ((e2e_batch, rows, rows), np.float32), ((e2e_batch, rows), np.int32)]
        else:
            shapes = [((e2e_batch, rows, rows), np.float32), ((e2e_batch, rows, cols), np.float32)]
        # one pinned block for all outputs (L | U | P, Q | R): the copy probe below returns exactly the same bytes
        pO, hO = self.pinned(out_bytes, np.uint8, (out_bytes,))
        outs, ptrs, off = [], [pA, pO], 0
        for shp, dt in shapes:
            nb = int(np.prod(shp)) * 4
            outs.append(hO[off:off + nb].view(dt).reshape(shp))
            off += nb
        assert off == out_bytes
        dcode = {**hd.CHOL_DESIGNS, **hd.LU_DESIGNS, **hd.QR_DESIGNS}[design]
        pcode = hd.PATHS[args.path] | (hd.FLAG_FUSED if args.fused else 0)

        def host_step():
            if kind == "chol":
                rc = lib.hlsd_cholesky_host(hA.ctypes.data, outs[0].ctypes.data, rows, e2e_batch, dcode, pcode)
            elif kind == "lu":
                rc = lib.hlsd_lu_host(hA.ctypes.data, outs[0].ctypes.data, outs[1].ctypes.data, outs[2].ctypes.data, rows,
                                      e2e_batch, dcode, pcode)
            else:
                rc = lib.hlsd_qr_host(hA.ctypes.data, outs[0].ctypes.data, outs[1].ctypes.data, rows, cols, e2e_batch, dcode, pcode)
            if rc:
                raise RuntimeError(lib.hlsd_last_error().decode())

        def copy_step():  # the same staging pipeline without the kernel: what the box's links allow
            rc = lib.hlsd_selftest_copy_host(hA.ctypes.data, hO.ctypes.data, in_mat, out_mat, e2e_batch)
            if rc:
                raise RuntimeError(lib.hlsd_last_error().decode())

        host_step()  # warm-up (stream/buffer creation paths, page faults)
        self.barrier()
        sampler.region(True)
        t0 = time.perf_counter()
        for _ in range(e2e_steps):
            host_step()  # returns when the outputs are complete in host memory
        e2e_s = time.perf_counter() - t0
        sampler.region(False)
        e2e_units, e2e_ms = combine_ranks(e2e_batch * e2e_steps, e2e_s * 1e3, self.dev)
        e2e = {"value": e2e_units / (e2e_ms * 1e-3), "unit": METRIC, "h2d_bytes_per_step": in_bytes,
               "d2h_bytes_per_step": out_bytes, "steps": e2e_steps, "ms_per_step": e2e_ms / e2e_steps,
               "batch_per_step": e2e_batch, "cpu_affinity": sampler.affinity,
               "api": f"hlsd_{'cholesky' if kind == 'chol' else kind}_host (C ABI, pinned host buffers, chunked 3-stream pipeline)"}
        if e2e_batch < batch:
            e2e["note"] = (f"host buffers capped at {args.e2e_max_gb:g} GiB per rank (--e2e-max-gb): {e2e_batch} of the {batch} "
                           "matrices per step; the rate is PCIe-bound and does not depend on the batch beyond a few chunks")
        # the e2e result must be the same answer the resident leg produced: every output, a sample across the batch
        idx = np.unique(np.linspace(0, e2e_batch - 1, 512).astype(np.int64))
        tidx = torch.from_numpy(idx).to(self.dev)
        res_outs = out if isinstance(out, tuple) else (out,)
        e2e["matches_resident"] = bool(all(np.array_equal(h[idx], d[tidx].cpu().numpy()) for h, d in zip(outs, res_outs)))
        e2e["matches_resident_sample"] = f"{len(idx)} matrices spread over the batch, every output array"
        # copy ceiling: all ranks copy at the same time, like the e2e leg (run last: it overwrites the outputs)
        copy_step()
        self.barrier()
        csteps = max(3, e2e_steps // 2)
        t0 = time.perf_counter()
        for _ in range(csteps):
            copy_step()
        c_s = time.perf_counter() - t0
        c_units, c_ms = combine_ranks(e2e_batch * csteps, c_s * 1e3, self.dev)
        ceiling = c_units / (c_ms * 1e-3)
        e2e["copy_ceiling"] = {"value": ceiling, "unit": METRIC, "h2d_gbs_this_rank": in_bytes * csteps / c_s / 1e9,
                               "d2h_gbs_this_rank": out_bytes * csteps / c_s / 1e9,
                               "how": "hlsd_selftest_copy_host: the same chunked 3-stream pipeline, host buffers and byte counts "
                                      "with the kernel launch left out, all ranks at once"}
        e2e["copy_ceiling_frac"] = e2e["value"] / ceiling if ceiling else None
        for p in ptrs:
            lib.hlsd_host_free(p)
        lib.hlsd_release_workspace()
        return e2e

    def e2e_single_process(self, name, e2e_steps):
        """Rank 0 alone drives ALL GPUs of the job through hlsd_*_host_multi (one host batch split inside the library)."""
        import numpy as np
        torch, hd, lib, args = self.torch, self.hd, self.lib, self.args
        kind, rows, cols, full_batch, design = WORKLOADS[name]
        if kind != "chol":
            return None
        total = full_batch * self.world if args.scaling == "weak" else full_batch
        res = None
        if self.rank == 0:
            nbytes = total * rows * rows * 4
            pA, hA = self.pinned(nbytes, np.float32, (total, rows, rows))
            pL, hL = self.pinned(nbytes, np.float32, (total, rows, rows))
            src = self.operand(kind, rows, cols, full_batch, 99)
            for k in range(0, total, full_batch):
                torch.from_numpy(hA[k:k + full_batch]).copy_(src[:min(full_batch, total - k)])
            del src
            dcode, pcode = hd.CHOL_DESIGNS[design], hd.PATHS[args.path]

            def multi_step():
                rc = lib.hlsd_cholesky_host_multi(hA.ctypes.data, hL.ctypes.data, rows, total, dcode, pcode, None, self.world)
                if rc:
                    raise RuntimeError(lib.hlsd_last_error().decode())

            multi_step()
            t0 = time.perf_counter()
            for _ in range(e2e_steps):
                multi_step()
            dt = time.perf_counter() - t0
            res = {"value": total * e2e_steps / dt, "unit": METRIC, "h2d_bytes_per_step": nbytes, "d2h_bytes_per_step": nbytes,
                   "steps": e2e_steps, "ms_per_step": 1e3 * dt / e2e_steps, "devices": self.world,
                   "api": "hlsd_cholesky_host_multi (one process, one host batch split over the device list inside the library)"}
            lib.hlsd_host_free(pA)
            lib.hlsd_host_free(pL)
            lib.hlsd_release_workspace()
        self.barrier()
        return res

    def per_n(self):
        """BASELINE.json's metric per N: resident, auto dispatch, single calls of ~3 GiB algorithmic traffic."""
        torch, hd = self.torch, self.hd
        peak_gbs = float(self.peaks.get("hbm_gbs", 6650.0))
        rows_out = []
        for kind in ("chol", "lu", "qr"):
            for n in (4, 8, 16, 32, 64, 128, 256, 512, 1024):
                if kind == "qr" and n > 128:
                    continue
                ab = algorithmic_bytes(kind, n, n)
                batch = int(min(8 << 20, max(64, (3 << 30) // ab)))
                if n >= 256:
                    batch = min(batch, 4096 * 1024 * 1024 // (n * n))
                A = self.operand(kind, n, n, batch, 7 + n)
                fn = {"chol": hd.cholesky, "lu": hd.lu, "qr": hd.qr}[kind]
                for _ in range(3):
                    fn(A)
                torch.cuda.synchronize()
                a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                reps = 5
                a.record()
                for _ in range(reps):
                    fn(A)
                b.record()
                torch.cuda.synchronize()
                ms = a.elapsed_time(b) / reps
                rows_out.append({"op": kind, "n": n, "batch": batch, "ms": ms, "fact_per_s": batch / (ms * 1e-3),
                                 "gflops": flops(kind, n, n) * batch / (ms * 1e-3) / 1e9,
                                 "gbs": ab * batch / (ms * 1e-3) / 1e9, "frac_of_hbm_peak": ab * batch / (ms * 1e-3) / 1e9 / peak_gbs,
                                 "path": "tensor (tolerance)" if (kind != "qr" and n >= 256) else "exact"})
                del A
                torch.cuda.empty_cache()
        return rows_out


# ------------------------------------------------------------------------------------------------
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default=HEADLINE, choices=sorted(WORKLOADS))
    ap.add_argument("--path", default="auto")
    ap.add_argument("--fused", action="store_true", help="HLSD_FLAG_FUSED (opt-in FMA contraction; not the parity configuration)")
    ap.add_argument("--scaling", default="weak", choices=["weak", "strong"])
    ap.add_argument("--batch", type=int, default=0, help="override the batch of --workload")
    ap.add_argument("--e2e-steps", type=int, default=10)
    ap.add_argument("--e2e-max-gb", type=float, default=16.0, help="cap on the pinned host buffers of one e2e leg, per rank")
    ap.add_argument("--cpu-seconds", type=float, default=12.0)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true", help="skip the host-buffer leg (exploration runs only)")
    ap.add_argument("--headline-only", action="store_true", help="skip the `workloads` and `per_n` sections")
    args = ap.parse_args()

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    kind, rows, cols, batch, design = WORKLOADS[args.workload]
    if args.batch:
        batch = args.batch
    config = workload_config(args.workload, args.path, args.scaling)
    config["batch"] = batch

    if args.impl == "reference":
        if rank != 0:
            return 0
        k = max(1, args.steps)
        cr = CpuReference(kind, rows, cols, design, target_s=min(args.cpu_seconds, 100.0 / (args.warmup + k)))
        for _ in range(args.warmup):
            cr.run_once()
        t_total = sum(cr.run_once() for _ in range(k))
        val = cr.count * k / t_total
        cb = cr.describe(val, t_total / k)
        print(json.dumps({"impl": "reference", "metric": METRIC, "value": val, "unit": METRIC, "n_gpus": args.gpus,
                          "steps": k, "warmup": args.warmup, "ms_per_step": 1e3 * t_total / k, "higher_is_better": True,
                          "scaling": args.scaling, "vs_baseline": None, "dtype": "f32", "data": "synthetic",
                          "config": config, "cpu_baseline": cb,
                          "e2e": {"value": val, "unit": METRIC, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}))
        return 0

    import torch
    import torch.distributed as dist

    if not torch.cuda.is_available():
        sys.exit("bench.py: no CUDA device; this library has no CPU path (use --impl reference for the CPU reference)")
    sampler = ClockSampler(local_rank)
    sampler.bind_cpu_near_gpu()  # before the CUDA context and the pinned allocations exist
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    runner = Runner(args, rank, world, local_rank)

    head = runner.run(args.workload, args.steps, args.warmup, args.e2e_steps, args.cpu_seconds, sampler)
    single = None
    if world > 1 and not args.no_e2e and args.workload == HEADLINE:
        single = runner.e2e_single_process(args.workload, max(3, args.e2e_steps // 2))
    extras, per_n = [], None
    if not args.headline_only and args.workload == HEADLINE and not args.batch:
        for name in EXTRA_WORKLOADS:
            try:
                r = runner.run(name, min(args.steps, 20 if WORKLOADS[name][1] < 256 else 10), 3, min(args.e2e_steps, 5),
                               min(args.cpu_seconds, 4.0), sampler)
            except Exception as exc:  # noqa: BLE001  (a secondary workload must not take the headline line down)
                r = {"name": name, "error": f"{type(exc).__name__}: {exc}"}
            extras.append(r)
        if world == 1:
            try:
                per_n = runner.per_n()
            except Exception as exc:  # noqa: BLE001
                per_n = [{"error": f"{type(exc).__name__}: {exc}"}]
    clocks = sampler.summary()

    if rank == 0:
        line = {"metric": METRIC, "value": head["value"], "unit": METRIC, "n_gpus": world, "steps": args.steps,
                "warmup": max(3, args.warmup), "ms_per_step": head["ms_per_step"], "higher_is_better": True,
                "scaling": args.scaling, "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": config,
                "roofline": head["roofline"], "e2e": head.get("e2e"), "gpu_launches": head["gpu_launches"], "clocks": clocks}
        if single is not None:
            line["e2e_single_process"] = single
        if "cpu_baseline" in head:
            line["cpu_baseline"] = head["cpu_baseline"]
        if extras:
            line["workloads"] = extras
        if per_n is not None:
            line["per_n"] = per_n
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
