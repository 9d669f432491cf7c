#!/usr/bin/env python
"""bench.py -- k-points/s of the k-grid evaluation hot path (Fe 18-WF AHC+DOS Fermi scan).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

A "step" is one pass of the hot path over one batch of K-blocks of the BASELINE config-2 grid
(bcc Fe, 18 WF, nR=95; 400^3 k-grid split as NKdiv=20 x NKFFT=20, 2000 Fermi levels over 12..22 eV):
`--blocks` K-blocks of 20^3 = 8000 k-points per GPU (weak scaling: per-GPU work is fixed; the K-block
list shards across ranks with no data-path collective; one all-reduce of the Fermi-scan arrays per step).
The system is the reference's test data (tests/golden/fe_system.npz, generated from
tests/reference/systems/Fe_W90 of the reference).

Printed JSON (one line, rank 0):
  value      whole-job k-points/s, dK list / weights / outputs resident in HBM
  e2e        the same through the public host API (Engine.scan: host buffers, H2D + D2H inside)
  roofline   dominant kernel, timed live with CUDA events inside the library (option "timing")
  cpu_baseline  the oracle (numpy restatement of the reference) on 1 host core, bounded sample
With `--impl reference` the CPU port runs on all host cores (multiprocessing over K-blocks).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "k-points/s, Fe 18-WF AHC+DOS Fermi scan"
UNIT = "k-points/s"
FE = os.path.join(ROOT, "tests", "golden", "fe_system.npz")
NKFFT = [20, 20, 20]
NKDIV = [20, 20, 20]
EFERMI = np.linspace(12.0, 22.0, 2000)
# algorithmic work per k-point: M = 10 matrices of 18^2 complex128
NW = 18
BYTES_PER_K = 16 * NW * NW * 10
FLOPS_PER_K = {"fourier": 5 * np.log2(8000) * NW * NW * 10, "eigh": 36 * NW ** 3, "rotate": 16 * NW ** 3 * 9 + 72 * NW * NW}
STAGES = ["fourier", "eigh", "rotate", "identity", "scan"]


def config(blocks, n_gpus):
    return {"workload": f"bcc Fe 18-WF (nR=95, tests-data system) AHC+DOS, K-blocks of NKFFT=20^3 from the 400^3 grid "
                        f"(NKdiv=20), 2000 Fermi levels 12..22 eV; step = {blocks} K-blocks x 8000 k per GPU",
            "blocks_per_gpu": blocks, "kpoints_per_step": blocks * 8000 * n_gpus, "nEF": len(EFERMI),
            "parallelism": f"K-block sharding x{n_gpus}, one all-reduce of the scan arrays per step",
            "l2": "per-step working set (X(k) records, >= 2 GB) exceeds the 126 MB L2; no explicit flush"}


# ------------------------------------------------------------------------------------------ clocks
class ClockSampler:
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.samples, self.proc = index, [], None

    # NVML bit masks of nvmlClocksEventReasons (nvml.h): the reasons the timing rules reject or note
    NVML_REASONS = {0x8: "hw_slowdown", 0x40: "hw_thermal_slowdown", 0x20: "sw_thermal_slowdown", 0x4: "sw_power_cap"}

    def _nvml_loop(self):
        import pynvml as nv
        h = nv.nvmlDeviceGetHandleByIndex(self.index)
        smax = nv.nvmlDeviceGetMaxClockInfo(h, nv.NVML_CLOCK_SM)
        get_reasons = getattr(nv, "nvmlDeviceGetCurrentClocksEventReasons", None) or nv.nvmlDeviceGetCurrentClocksThrottleReasons
        while not self.stop_flag.is_set():
            mask = int(get_reasons(h))
            flags = ["Active" if mask & bit else "Not Active" for bit in (0x8, 0x40, 0x20, 0x4)]
            self.samples.append(",".join([str(nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM)), str(smax)] + flags))
            time.sleep(0.005)

    def start(self):
        # the timed region is ~0.1 s: NVML is polled every 5 ms from a thread (the same counters nvidia-smi prints);
        # `nvidia-smi -lms` (one sample per 100 ms after its start-up) is the fallback
        try:
            import pynvml as nv
            nv.nvmlInit()
            self.stop_flag = threading.Event()
            self.t = threading.Thread(target=self._nvml_loop, daemon=True)
            self.t.start()
            self.proc = "nvml"
            return
        except Exception:
            self.proc = None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.samples.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        if self.proc == "nvml":
            self.stop_flag.set()
        else:
            self.proc.terminate()
        self.t.join(timeout=2)
        sm, smax, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for s in self.samples:
            f = [x.strip() for x in s.split(",")]
            try:
                sm.append(float(f[0]))
                smax.append(float(f[1]))
            except (ValueError, IndexError):
                continue
            for n, v in zip(names, f[2:6]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(smax) if smax else None,
                "reasons": sorted(reasons), "samples": len(sm), "source": "nvml" if self.proc == "nvml" else "nvidia-smi"}


# ------------------------------------------------------------------------------------------ CPU arms
def _oracle_block(args):
    """One K-block through the oracle (CPU port of the reference algorithm)."""
    dK, nkfft = args
    os.environ.setdefault("OMP_NUM_THREADS", "1")
    from oracle import wb_oracle as orc
    osys = _oracle_block.sys
    data = orc.OracleDataK(osys, dK, nkfft)
    orc.AHC(data, EFERMI)
    orc.DOS(data, EFERMI)
    return data.nk


def _oracle_init():
    for v in ("OMP_NUM_THREADS", "OPENBLAS_NUM_THREADS", "MKL_NUM_THREADS"):
        os.environ[v] = "1"
    from oracle import wb_oracle as orc
    _oracle_block.sys = orc.OracleSystem.from_npz(FE)


def cpu_baseline_1core(nkfft=(20, 20, 20)):
    """oracle on 1 core, one K-block of the same workload (same system, same 2000 Fermi levels)."""
    _oracle_init()
    t0 = time.perf_counter()
    nk = _oracle_block((np.array([1 / 400, 2 / 400, 3 / 400]), list(nkfft)))
    dt = time.perf_counter() - t0
    return {"value": nk / dt, "unit": UNIT, "cores": 1, "kind": "port",
            "sample": f"1 K-block of NKFFT={list(nkfft)} ({nk} k-points) of the same workload, {dt:.1f} s; "
                      "oracle/wb_oracle.py (numpy restatement of the reference, per-k Python loops as in the reference)"}


def run_reference_arm(args):
    """`--impl reference`: the CPU port on all host cores, on the SAME K-blocks the GPU arm evaluates (NKFFT = 20^3
    sub-grids at the shifts of the 400^3 grid's K-list, same calculators and Fermi levels); each step = one K-block
    per core (the reference parallelises over K-blocks, run_grid.py:258-265)."""
    import multiprocessing as mp
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import wannierberri_b200 as wb   # host-side grid only (K-list); no GPU work in this arm
    cores = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
    shifts, factors = wb.Grid(wb.System_R.from_npz(FE), NKdiv=NKDIV, NKFFT=NKFFT).K_arrays()
    cursor = [0]
    with mp.Pool(cores, initializer=_oracle_init) as pool:
        def step():
            idx = (cursor[0] + np.arange(cores)) % len(factors)   # the GPU arm's K-list, consecutive K-blocks
            cursor[0] += cores
            return sum(pool.map(_oracle_block, [(shifts[i], NKFFT) for i in idx], chunksize=1))
        for _ in range(args.warmup):
            step()
        t0 = time.perf_counter()
        nk = 0
        for _ in range(args.steps):
            nk += step()
        dt = time.perf_counter() - t0
    value = nk / dt
    sample = (f"{cores} K-blocks of NKFFT={NKFFT} ({int(np.prod(NKFFT))} k-points each) per step, consecutive entries of the "
              f"400^3 grid's K-list (the GPU arm's list), one per core: multiprocessing.Pool({cores}), BLAS threads = 1")
    cfg = config(args.blocks, args.gpus)
    cfg["reference_arm_step"] = {"blocks_per_step": cores, "NKFFT": NKFFT, "kpoints_per_step": cores * int(np.prod(NKFFT))}
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "tests-data system (Fe_W90 of the reference); K-block shifts of the 400^3 grid",
        "config": cfg,
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0}))


# ------------------------------------------------------------------------------------------ GPU arm
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200")
    ap.add_argument("--blocks", type=int, default=32, help="K-blocks (of 8000 k-points) per GPU per step")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference_arm(args)

    import torch
    import torch.distributed as dist
    import wannierberri_b200 as wb
    from wannierberri_b200 import _lib
    import ctypes as C

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise RuntimeError("bench.py needs a CUDA device (no CPU fallback)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        # NCCL prints its version banner on stdout at the first collective: keep stdout for the ONE JSON line
        sys.stdout.flush()
        saved = os.dup(1)
        os.dup2(2, 1)
        try:
            dist.init_process_group("nccl", device_id=dev)
            t = torch.zeros(1, device=dev)
            dist.all_reduce(t)
            torch.cuda.synchronize()
        finally:
            sys.stdout.flush()
            os.dup2(saved, 1)
            os.close(saved)

    system = wb.System_R.from_npz(FE)
    st = wb.calculators.static
    calcs = dict(ahc=st.AHC(Efermi=EFERMI), dos=st.DOS(Efermi=EFERMI))
    specs = [s for c in calcs.values() for s in c.specs()]
    eng = wb.Engine(system, device=local)
    eng.plan(NKFFT, [s.formula for s in specs], external_terms=True)

    # this rank's K-blocks: a contiguous chunk of the 400^3 grid's K-list (shifts of the 20^3 sub-grid)
    grid = wb.Grid(system, NKdiv=NKDIV, NKFFT=NKFFT)
    shifts, factors = grid.K_arrays()
    nb = args.blocks
    lo = (rank * nb) % len(factors)
    idx = (lo + np.arange(nb)) % len(factors)
    dK_h = np.ascontiguousarray(shifts[idx])
    w_h = np.ascontiguousarray(factors[idx])
    dK_pin = torch.from_numpy(dK_h).pin_memory()
    w_pin = torch.from_numpy(w_h).pin_memory()
    dK_d, w_d = dK_pin.to(dev), w_pin.to(dev)
    nout = sum(s.size for s in specs)
    out_d = torch.zeros(nout, dtype=torch.float64, device=dev)
    kpts_step = nb * int(np.prod(NKFFT)) * world

    def step_dev():
        eng.scan_dev(dK_d, w_d, specs, out_d)
        if world > 1:
            dist.all_reduce(out_d)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier()
        e0.record()
        for _ in range(steps):
            fn()
        e1.record()
        barrier()
        ms = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms.item())

    for _ in range(max(args.warmup, 3)):
        step_dev()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    l0 = eng.kernel_launches
    ms = timed(step_dev, args.steps)
    launches = eng.kernel_launches - l0
    clocks = sampler.stop() if rank == 0 else None
    value = kpts_step * args.steps / (ms * 1e-3)

    # ---- e2e: public host API, host buffers in and out every step
    host_out = {}

    def step_e2e():
        arrays = eng.scan(dK_h, w_h, specs)
        if world > 1:
            t = torch.from_numpy(np.concatenate([a.ravel() for a in arrays])).to(dev)
            dist.all_reduce(t)
            arrays = [t.cpu().numpy()]
        host_out["last"] = arrays

    step_e2e()
    ms_e2e = timed(step_e2e, args.steps)
    e2e = {"value": kpts_step * args.steps / (ms_e2e * 1e-3), "unit": UNIT,
           "h2d_bytes_per_step": int(dK_h.nbytes + w_h.nbytes), "d2h_bytes_per_step": int(nout * 8),
           "api": "wannierberri_b200.Engine.scan -> wbgpu_static_scan (host pointers)"}

    # ---- per-stage device times (CUDA events inside the library, on its stream)
    eng.set_option("timing", 1)
    for _ in range(args.steps):
        step_dev()
    torch.cuda.synchronize()
    ms_st = (C.c_double * 5)()
    calls = (C.c_int64 * 5)()
    _lib.check(_lib.lib().wbgpu_stage_times(eng._ctx, ms_st, calls))
    eng.set_option("timing", 0)
    stage_ms = {STAGES[i]: ms_st[i] / args.steps for i in range(5)}
    stage_calls = {STAGES[i]: int(calls[i]) // args.steps for i in range(5)}
    dominant = max(("fourier", "eigh", "rotate"), key=lambda s: stage_ms[s])

    if rank == 0:
        fp64 = {}
        for kind, name in ((0, "dfma"), (1, "dmma")):
            t = C.c_double()
            _lib.check(_lib.lib().wbgpu_fp64_peak(local, kind, C.byref(t)))
            fp64[name] = t.value
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except OSError:
            pass
        hbm_peak = peaks.get("hbm_gbs", 6650.0)
        kpts_rank = nb * int(np.prod(NKFFT))
        k_per_launch = kpts_rank / max(stage_calls[dominant], 1)
        dur = stage_ms[dominant] / max(stage_calls[dominant], 1) * 1e-3
        achieved = FLOPS_PER_K[dominant] * k_per_launch / dur / 1e12
        peak64 = max(fp64.values())
        traffic = None
        try:  # dram bytes of the dominant kernel from the committed ncu --set full capture, scaled to this launch
            tj = json.load(open(os.path.join(ROOT, "profiles", "r1_traffic.json")))
            if dominant == "rotate":
                traffic = tj["bytes_per_kpoint"] * k_per_launch
        except (OSError, KeyError):
            pass
        roofline = {"kernel": dominant, "bound": "tensor", "pipe": "FP64 tensor (mma.sync.m8n8k4.f64)", "achieved": achieved,
                    "peak": peak64, "unit": "TFLOP/s", "frac": achieved / peak64, "traffic": traffic,
                    "traffic_source": "profiles/r1_traffic.json (ncu dram__bytes_read.sum + dram__bytes_write.sum per k-point x k-points per launch)",
                    "peak_source": f"measured in this run: DFMA {fp64['dfma']:.1f}, DMMA(m8n8k4) {fp64['dmma']:.1f} TFLOP/s",
                    "algorithmic_flops_per_k": FLOPS_PER_K[dominant], "avg_launch_ms": dur * 1e3,
                    "kpoints_per_launch": k_per_launch}
        step_s = sum(stage_ms.values()) * 1e-3
        roofline_hbm = {"bound": "hbm", "achieved": BYTES_PER_K * kpts_rank / step_s / 1e9, "peak": hbm_peak, "unit": "GB/s",
                        "frac": BYTES_PER_K * kpts_rank / step_s / 1e9 / hbm_peak,
                        "peak_source": "MEASURED_PEAKS.json" if peaks else "fallback (B200_PROFILING.md)",
                        "note": "whole pipeline, algorithmic X(k) bytes (16*nw^2*10 per k-point) / sum of stage times"}
        cpu = None if (args.no_cpu_baseline or world > 1) else cpu_baseline_1core()   # rank 0 at N = 1 only
        print(json.dumps({
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f64",
            "data": "tests-data system (Fe_W90 of the reference, 18 WF, nR=95); K-block shifts of the 400^3 grid",
            "config": config(nb, world), "e2e": e2e, "gpu_launches": int(launches), "clocks": clocks,
            "roofline": roofline, "roofline_hbm": roofline_hbm, "stage_ms_per_step": stage_ms,
            "fp64_peak_tflops": fp64, "cpu_baseline": cpu}))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
