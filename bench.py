#!/usr/bin/env python3
"""Benchmark of the hot path BASELINE.json names: batched factorisations per second.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload chol16] [--impl ours|reference]

Default workload (N=1): BASELINE.json configs[1] — batched Cholesky fp32 N=16, batch 1M synthetic
SPD matrices (genA.m formula) on one B200.  One "step" = one pass of the factorisation over the
whole batch.  With N>1 (torchrun, one rank per GPU) every rank factors its own batch of the same
size (weak scaling, plain batch splitting, no collective on the data path; NCCL is used only for
the barrier and to combine the per-rank counters).

One JSON line on stdout (rank 0):
  value        factorizations/s, inputs resident in HBM, device-timed (CUDA events), max over ranks
  e2e          same metric through the C ABI's host-pointer entry point (hlsd_cholesky_host) with
               pinned HOST buffers: H2D and D2H copies inside the timed region
  roofline     algorithmic HBM bytes of the kernel / its measured duration, vs MEASURED_PEAKS.json
  cpu_baseline the reference's own C++ (oracle/_ref, g++ + HLS header stand-ins) on the host cores
`--impl reference` times only that CPU reference (rank 0) and prints the same line shape.
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
    # name: (kind, rows, cols, batch per GPU, design, reference design tag)
    "chol16": ("chol", 16, 16, 1 << 20, "cholesky_v1.3"),
    "lu64": ("lu", 64, 64, 262144, "lu_v1.1"),
    # the same elimination with the pivot search switched off (P[k] = k): BASELINE.json words its LU config as
    # "unpivoted"; the reference designs themselves pivot, so lu64 above is the parity workload
    "lu64np": ("lu", 64, 64, 262144, "nopivot"),
    "qr128": ("qr", 128, 128, 16384, "qr_v2.1"),
    "chol64": ("chol", 64, 64, 262144, "cholesky_v1.3"),
    "chol1024": ("chol", 1024, 1024, 4096, "cholesky_v1.3"),
    "lu1024": ("lu", 1024, 1024, 4096, "lu_v1.1"),
}


# ------------------------------------------------------------------------------------------------
# multi-GPU plan: plain batch splitting (also exercised on CPU by tests/test_sharding.py)
# ------------------------------------------------------------------------------------------------
def shard_batch(batch, rank, world):
    """Contiguous, balanced [start, end) span of `batch` units for `rank` of `world`."""
    base, extra = divmod(batch, world)
    start = rank * base + min(rank, extra)
    return start, start + base + (1 if rank < extra else 0)


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


# ------------------------------------------------------------------------------------------------
# clocks: NVML poller running during the timed regions
# ------------------------------------------------------------------------------------------------
class ClockSampler:
    REASONS = {0x4: "sw_power_cap", 0x8: "hw_slowdown", 0x20: "sw_thermal_slowdown", 0x40: "hw_thermal_slowdown",
               0x80: "hw_power_brake_slowdown", 0x2: "applications_clocks_setting", 0x10: "sync_boost"}

    def __init__(self, index):
        self.samples, self.reasons, self.max_mhz, self._stop, self._on = [], set(), None, False, False
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv, self.h = pynvml, pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
            self.thread = threading.Thread(target=self._run, daemon=True)
            self.thread.start()
        except Exception as exc:  # noqa: BLE001
            self.nv, self.err = None, str(exc)

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
        probe = max(self.cores * 4, 64)
        while True:  # grow the probe until it runs long enough to extrapolate from
            self.A = self._make(probe)
            dt = max(self.run_once(), 1e-6)
            if dt >= 0.25 or probe >= (1 << 20):
                break
            probe *= 4
        count = int(min(max(64, target_s * probe / dt), 1 << 21))
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


# ------------------------------------------------------------------------------------------------
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="chol16", choices=sorted(WORKLOADS))
    ap.add_argument("--path", default="auto")
    ap.add_argument("--batch", type=int, default=0, help="override the per-GPU batch")
    ap.add_argument("--e2e-steps", type=int, default=3)
    ap.add_argument("--cpu-seconds", type=float, default=12.0)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true", help="skip the host-buffer leg (exploration runs only)")
    args = ap.parse_args()

    kind, rows, cols, batch, design = WORKLOADS[args.workload]
    if args.batch:
        batch = args.batch
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    workload_name = f"batched {kind} fp32 {rows}x{cols}, batch {batch} per GPU, design {design}, synthetic genA operands"
    config = {"workload": workload_name, "batch_per_gpu": batch, "rows": rows, "cols": cols, "design": design,
              "kernel_path": args.path, "l2": "inputs larger than L2 (no flush needed)"}

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
        config = dict(config, batch_per_step=cr.count)
        print(json.dumps({"impl": "reference", "metric": METRIC, "value": val, "unit": METRIC, "n_gpus": args.gpus,
                          "steps": k, "warmup": args.warmup, "ms_per_step": 1e3 * t_total / k, "higher_is_better": True,
                          "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
                          "config": config, "cpu_baseline": cb,
                          "e2e": {"value": val, "unit": METRIC, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}))
        return 0

    import numpy as np
    import torch
    import torch.distributed as dist
    import hls_designs_b200 as hd

    if not torch.cuda.is_available():
        sys.exit("bench.py: no CUDA device; this library has no CPU path (use --impl reference for the CPU reference)")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    lib = hd.load()
    lib.hlsd_set_device(local_rank)
    sampler = ClockSampler(local_rank)

    # ---- synthetic operands, generated on the device for the resident leg ----
    g = torch.Generator(device=dev).manual_seed(1234 + rank)
    if kind == "chol":
        A = torch.rand((batch, rows, rows), generator=g, device=dev)
        A = 0.5 * (A + A.transpose(1, 2))
        A.diagonal(dim1=1, dim2=2).add_(float(rows))
        A = A.contiguous()
    else:
        A = torch.randint(1, 101, (batch, rows, cols), generator=g, device=dev).float()
    fn = {"chol": hd.cholesky, "lu": hd.lu, "qr": hd.qr}[kind]

    def step():
        return fn(A, design=design, path=args.path)

    for _ in range(max(3, args.warmup)):
        out = step()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    stream = torch.cuda.current_stream()
    evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    launches0 = hd.launch_count()
    sampler.region(True)
    t_begin = torch.cuda.Event(enable_timing=True)
    t_end = torch.cuda.Event(enable_timing=True)
    t_begin.record(stream)
    for s_ev, e_ev in evs:
        s_ev.record(stream)
        out = step()
        e_ev.record(stream)
    t_end.record(stream)
    torch.cuda.synchronize()
    sampler.region(False)
    if world > 1:
        dist.barrier()
    launches = hd.launch_count() - launches0
    total_ms = t_begin.elapsed_time(t_end)
    kern_ms = [s.elapsed_time(e) for s, e in evs]
    units, worst_ms = combine_ranks(batch * args.steps, total_ms, dev)
    value = units / (worst_ms * 1e-3)

    # ---- roofline of the dominant kernel (the step is one or a few launches of it) ----
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:  # noqa: BLE001
        pass
    peak_gbs = float(peaks.get("hbm_gbs", 6650.0))
    peak_src = "MEASURED_PEAKS.json hbm_gbs (measured)" if "hbm_gbs" in peaks else "fallback 6650 GB/s (B200_PROFILING.md)"
    med_ms = statistics.median(kern_ms)
    bytes_per_launch = algorithmic_bytes(kind, rows, cols) * batch
    achieved = bytes_per_launch / (med_ms * 1e-3) / 1e9
    traffic = None
    try:  # dram__bytes_read.sum + dram__bytes_write.sum of the dominant kernel, from the committed ncu capture
        traffic = json.load(open(os.path.join(ROOT, "profiles", "traffic.json"))).get(args.workload)
    except Exception:  # noqa: BLE001
        pass
    gflops = flops(kind, rows, cols) * batch / (med_ms * 1e-3) / 1e9
    roofline = {"bound": "hbm", "achieved": achieved, "peak": peak_gbs, "unit": "GB/s", "frac": achieved / peak_gbs,
                "traffic": traffic if batch == WORKLOADS[args.workload][3] else None, "peak_source": peak_src,
                "kernel_ms_median": med_ms, "kernel_ms_min": min(kern_ms), "kernel_ms_max": max(kern_ms),
                "algorithmic_bytes_per_launch": bytes_per_launch, "gflops": gflops}
    if kind in ("chol", "lu") and rows >= 256:
        # blocked tensor-core path: the step is ~25 launches; the dominant kernel (the tcgen05 tile kernel) does the
        # useful flops (N^3/3 Cholesky, 2N^3/3 LU) three times (3xTF32).  Peak: tf32 dense = half the measured bf16 rate.
        peak_tf = float(peaks.get("bf16_tflops_sustained", 1400.0)) / 2.0
        ach = gflops / 1e3
        roofline = {"bound": "tensor", "achieved": ach, "peak": peak_tf, "unit": "TFLOP/s", "frac": ach / peak_tf,
                    "traffic": None, "kernel_ms_median": med_ms, "kernel_ms_min": min(kern_ms),
                    "kernel_ms_max": max(kern_ms), "gflops": gflops,
                    "peak_source": "MEASURED_PEAKS.json bf16_tflops_sustained / 2 (tf32 runs at half the bf16 rate)",
                    "note": "achieved = useful flops of the whole step (N^3/3 Cholesky, 2N^3/3 LU) / step time; the tensor "
                            "pipe executes 3x that (3xTF32); a step is one launch per panel and stage, see profiles/"}

    e2e = None
    if not args.no_e2e:
        # ---- end to end through the host-pointer C entry point, pinned host buffers ----
        def pinned(nbytes, dtype, shape):
            p = lib.hlsd_host_alloc(nbytes)
            if not p:
                raise MemoryError(lib.hlsd_last_error().decode())
            arr = np.ctypeslib.as_array((ctypes.c_byte * nbytes).from_address(p)).view(dtype).reshape(shape)
            return p, arr

        e2e_batch = batch
        in_bytes = e2e_batch * rows * cols * 4
        pA, hA = pinned(in_bytes, np.float32, (e2e_batch, rows, cols))
        hA[...] = A.cpu().numpy()
        outs, ptrs = [], [pA]
        if kind == "chol":
            shapes = [((e2e_batch, rows, rows), np.float32)]
        elif kind == "lu":
            shapes = [((e2e_batch, rows, rows), np.float32), ((e2e_batch, rows, rows), np.float32), ((e2e_batch, rows), np.int32)]
        else:
            shapes = [((e2e_batch, rows, rows), np.float32), ((e2e_batch, rows, cols), np.float32)]
        out_bytes = 0
        for shp, dt in shapes:
            nb = int(np.prod(shp)) * 4
            p, arr = pinned(nb, dt, shp)
            ptrs.append(p)
            outs.append(arr)
            out_bytes += nb
        dcode = {**hd.CHOL_DESIGNS, **hd.LU_DESIGNS, **hd.QR_DESIGNS}[design]
        pcode = hd.PATHS[args.path]

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

        host_step()  # warm-up (stream/buffer creation paths, page faults)
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        sampler.region(True)
        t0 = time.perf_counter()
        for _ in range(args.e2e_steps):
            host_step()  # returns when the outputs are complete in host memory
        e2e_s = time.perf_counter() - t0
        sampler.region(False)
        e2e_units, e2e_ms = combine_ranks(e2e_batch * args.e2e_steps, e2e_s * 1e3, dev)
        e2e = {"value": e2e_units / (e2e_ms * 1e-3), "unit": METRIC, "h2d_bytes_per_step": in_bytes,
               "d2h_bytes_per_step": out_bytes, "steps": args.e2e_steps, "ms_per_step": e2e_ms / args.e2e_steps,
               "api": f"hlsd_{'cholesky' if kind == 'chol' else kind}_host (C ABI, pinned host buffers, chunked 3-stream pipeline)"}
        # the e2e result must be the same answer the resident leg produced
        ref_out = out[0] if isinstance(out, tuple) else out
        first = outs[0][:256]
        e2e["matches_resident"] = bool(np.array_equal(first, ref_out[:256].cpu().numpy()))
        for p in ptrs:
            lib.hlsd_host_free(p)

    clocks = sampler.summary()
    if roofline["bound"] == "hbm":
        # context for the exact kernels that are not HBM-bound: the reference's arithmetic has no fused
        # multiply-add (every float operation is rounded separately), so the arithmetic ceiling is one
        # operation per FP32 lane and clock, not two
        props = torch.cuda.get_device_properties(dev)
        ghz = (clocks.get("sm_max_mhz") or 1965) / 1e3
        peak = props.multi_processor_count * 128 * ghz / 1e3  # Tflop/s without FMA
        roofline["fp32_unfused"] = {"achieved": gflops / 1e3, "peak": peak, "unit": "TFLOP/s", "frac": gflops / 1e3 / peak,
                                    "note": "useful flops of the factorisation / (SMs x 128 FP32 lanes x max clock); "
                                            "the exact kernels issue separately rounded multiplies and adds"}
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        cpu = cpu_reference_rate(kind, rows, cols, design, args.cpu_seconds)

    if rank == 0:
        line = {"metric": METRIC, "value": value, "unit": METRIC, "n_gpus": world, "steps": args.steps,
                "warmup": max(3, args.warmup), "ms_per_step": worst_ms / args.steps, "higher_is_better": True,
                "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": config,
                "roofline": roofline, "e2e": e2e, "gpu_launches": int(launches), "clocks": clocks}
        if cpu is not None:
            line["cpu_baseline"] = cpu
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
