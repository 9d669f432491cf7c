#!/usr/bin/env python
"""bench.py -- k-points/s of the k-grid evaluation hot path.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--config headline|1|2|3|4|5]
                    [--scaling weak|strong]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

Default (`--config headline`, BASELINE.json's metric): bcc Fe, 18 WF, AHC + DOS, K-blocks of NKFFT = 20^3 of the 400^3
grid (NKdiv = 20, the split of BASELINE config 2), 2000 Fermi levels 12..22 eV.  A "step" is one pass of the hot path over
`--blocks` K-blocks per GPU (weak scaling: per-GPU work fixed; the K-block list shards across ranks with no data-path
collective; one all-reduce of the scan arrays per step).  `--config 1..5` times the five BASELINE configurations as
written, each with the roofline of its dominant stage and a CPU leg; `--scaling strong` runs
the WHOLE config-2 grid (8000 K-blocks) through `run()` under torchrun, total work fixed.

Printed JSON (one line, rank 0):
  value         whole-job k-points/s, dK list / weights / outputs resident in HBM
  e2e           the same through the public host API (host buffers in and out, H2D + D2H inside the timed region)
  roofline      dominant stage, timed live with CUDA events inside the library (option "timing")
  cpu_baseline  the oracle (numpy restatement of the reference) on 1 host core, bounded sample of the same workload
With `--impl reference` the CPU port runs on all host cores (one K-block per core per step)."""
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

UNIT = "k-points/s"
GOLDEN = os.path.join(ROOT, "tests", "golden")
STAGES = ["fourier", "eigh", "rotate", "identity", "scan"]


# ------------------------------------------------------------------------------------------ workloads
def _shortest_R(count):
    """whole shells of the shortest lattice vectors of a cubic lattice, closed under R -> -R (config 5)"""
    rr = np.arange(-10, 11)
    allR = np.array([[x, y, z] for x in rr for y in rr for z in rr], dtype=int)
    n2 = (allR ** 2).sum(axis=1)
    cut = np.sort(n2)[count]
    return allR[n2 < cut], n2[n2 < cut]


def build_system(name):
    import wannierberri_b200 as wb
    if name == "fe":
        return wb.System_R.from_npz(os.path.join(GOLDEN, "fe_system.npz"))
    if name == "te":
        return wb.System_R.from_npz(os.path.join(GOLDEN, "te_system.npz"))
    if name == "synth32":
        return wb.synthetic_system(32, rmax=2, seed=20261017)
    if name == "synth128":
        rng = np.random.default_rng(20261017)
        iRvec, n2 = _shortest_R(4000)
        index = {tuple(R): i for i, R in enumerate(iRvec)}
        minus = np.array([index[tuple(-R)] for R in iRvec])
        nR, nw = len(iRvec), 128
        H = (rng.standard_normal((nR, nw, nw)) + 1j * rng.standard_normal((nR, nw, nw))) * np.exp(-np.sqrt(n2))[:, None, None]
        H = 0.5 * (H + H[minus].transpose(0, 2, 1).conj())
        s = wb.System_R(np.eye(3) * 4.0, iRvec, rng.random((nw, 3)) * 4.0)
        s.set_R_mat("Ham", H)
        return s
    raise ValueError(name)


def oracle_system(name):
    """the same model for the CPU oracle (oracle/wb_oracle.py: test infrastructure, CPU legs only)"""
    from oracle import wb_oracle as orc
    if name in ("fe", "te"):
        return orc.OracleSystem.from_npz(os.path.join(GOLDEN, name + "_system.npz"))
    s = build_system(name)
    return orc.OracleSystem(s.rvec.iRvec, s.real_lattice, s.wannier_centers_cart, dict(s._XX_R))


class Workload:
    """One benchmark configuration: system, calculators, K-block split, algorithmic work per k-point."""

    def __init__(self, key, title, system, NKdiv, NKFFT, blocks, static=None, kubo=None, cpu_nkfft=None, nmat=10, rot_mats=9,
                 formula_flops=0.):
        self.key, self.title, self.system = key, title, system
        self.NKdiv, self.NKFFT, self.blocks = list(NKdiv), list(NKFFT), blocks
        self.static = static or {}      # key -> (calculator name, Efermi, kwargs)
        self.kubo = kubo                # (name, Efermi, omega, kwargs)
        self.cpu_nkfft = list(cpu_nkfft or NKFFT)
        self.nmat, self.rot_mats, self.formula_flops = nmat, rot_mats, formula_flops

    @property
    def nk_block(self):
        return int(np.prod(self.NKFFT))

    def metric(self):
        return "k-points/s, Fe 18-WF AHC+DOS Fermi scan" if self.key == "headline" else f"k-points/s, {self.title}"

    def flops_per_k(self, nw):
        """R->k as an FFT, eigh 36 nw^3, rotations 16 nw^3 per matrix, formula"""
        return {"fourier": 5 * np.log2(self.nk_block) * nw * nw * self.nmat, "eigh": 36. * nw ** 3,
                "rotate": 16. * nw ** 3 * self.rot_mats + self.formula_flops, "scan": self.scan_flops(nw)}

    pairs_per_k = None   # Kubo: band pairs whose Fermi-factor difference is non-zero somewhere on the Efermi axis (probed)

    def scan_flops(self, nw):
        """Kubo accumulation in the kernel's own (difference) form: per frequency and contributing band pair one complex
        frequency factor (~30 flops) and 9 complex multiply-adds at each end of the pair's Efermi interval (2 x 9 x 8
        flops).  The reference's dense [n_omega x npair] @ [npair x nEF 9] contraction (dynamic.py:79-100) would be
        8 n_omega npair nEF 9 flops -- nEF / 2 times more; it is reported as `reference_dense_flops_per_k`."""
        if self.kubo is None:
            return 0.
        _, Ef, om, _ = self.kubo
        pairs = self.pairs_per_k if self.pairs_per_k is not None else nw * (nw - 1) / 2
        return len(om) * pairs * (30. + 2 * 9 * 8)

    def scan_dense_flops(self, nw):
        if self.kubo is None:
            return 0.
        _, Ef, om, _ = self.kubo
        return 8. * len(om) * (nw * (nw - 1) / 2) * len(Ef) * 9

    def bytes_per_k(self, nw):
        return 16. * nw * nw * self.nmat


def workloads(key):
    Ef2000 = np.linspace(12.0, 22.0, 2000)
    if key == "headline":
        return Workload("headline", "bcc Fe 18-WF AHC+DOS", "fe", [20] * 3, [20] * 3, 32,
                        static=dict(ahc=("AHC", Ef2000, {}), dos=("DOS", Ef2000, {})), nmat=10, rot_mats=9, formula_flops=72 * 18 * 18)
    if key == "1":
        Ef = np.linspace(12.0, 22.0, 1001)
        return Workload("1", "BASELINE config 1: bcc Fe 18-WF AHC+DOS, 48^3 grid (NKdiv 4 x NKFFT 12), 1001 Fermi levels, no symmetry",
                        "fe", [4] * 3, [12] * 3, 64, static=dict(ahc=("AHC", Ef, {}), dos=("DOS", Ef, {})), nmat=10, rot_mats=9,
                        formula_flops=72 * 18 * 18)
    if key == "2":
        return Workload("2", "BASELINE config 2: bcc Fe 18-WF AHC + orbital moment, 400^3 grid (NKdiv 20 x NKFFT 20), 2000 Fermi levels",
                        "fe", [20] * 3, [20] * 3, 32, static=dict(ahc=("AHC", Ef2000, {}), morb=("Morb", Ef2000, {})), cpu_nkfft=[12] * 3,
                        nmat=16, rot_mats=15, formula_flops=2 * 72 * 18 * 18)
    if key == "3":
        Ef = np.linspace(4.0, 8.0, 401)
        names = ("BerryDipole_FermiSurf", "GME_orb_FermiSurf", "GME_spin_FermiSurf")
        return Workload("3", "BASELINE config 3: Te 24-WF spinor BerryDipole + GME (orb, spin) Fermi-surface terms, 200^3 grid "
                             "(NKdiv 10 x NKFFT 20), 401 Fermi levels", "te", [10] * 3, [20] * 3, 16,
                        static={n: (n, Ef, {}) for n in names}, cpu_nkfft=[10] * 3, nmat=19, rot_mats=18, formula_flops=3 * 72 * 24 * 24)
    if key == "4":
        Ef, om = np.linspace(-1, 1, 200), np.linspace(0, 5, 500)
        return Workload("4", "BASELINE config 4: 32-WF Kubo optical conductivity, 500 frequencies x 200 Fermi levels, 128^3 grid "
                             "(NKdiv 8 x NKFFT 16)", "synth32", [8] * 3, [16] * 3, 4,
                        kubo=("OpticalConductivity", Ef, om, dict(smr_fixed_width=0.1)), cpu_nkfft=[5] * 3, nmat=7, rot_mats=6)
    if key == "5":
        Ef = np.linspace(-4, 4, 401)
        return Workload("5", "BASELINE config 5: synthetic 128-WF x 3959-R model, eigenvalues + d_aH rotations (DOS, CumDOS, "
                             "Ohmic_FermiSurf), 512^3 grid (NKdiv 16 x NKFFT 32)", "synth128", [16] * 3, [32] * 3, 1,
                        static=dict(dos=("DOS", Ef, {}), cumdos=("CumDOS", Ef, {}), ohmic=("Ohmic_FermiSurf", Ef, {})),
                        cpu_nkfft=[4] * 3, nmat=4, rot_mats=3)
    raise ValueError(f"unknown --config {key}")


def config_dict(W, blocks, n_gpus, scaling="weak"):
    nk = blocks * W.nk_block
    d = {"workload": f"{W.title}; K-blocks of NKFFT={W.NKFFT} at the shifts of the NKdiv={W.NKdiv} K-list; "
                     f"step = {blocks} K-blocks x {W.nk_block} k per GPU",
         "blocks_per_gpu": blocks, "kpoints_per_step": nk * n_gpus,
         "parallelism": f"K-block sharding x{n_gpus}, one all-reduce of the scan arrays per step",
         "l2": "per-step working set (X(k) records of the batch, >= 1 GB) exceeds the 126 MB L2; no explicit flush"}
    if W.static:
        d["nEF"] = int(len(next(iter(W.static.values()))[1]))
    if W.kubo:
        d["nEF"], d["nomega"] = int(len(W.kubo[1])), int(len(W.kubo[2]))
    if scaling == "strong":
        d["workload"] = f"{W.title}; the WHOLE grid ({int(np.prod(W.NKdiv))} K-blocks of NKFFT={W.NKFFT}) through run(), sharded over the ranks"
        d["parallelism"] = f"K-block sharding x{n_gpus} inside run(), one all-reduce"
    return d


# ------------------------------------------------------------------------------------------ clocks
class ClockSampler:
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.samples, self.proc = index, [], None

    def _nvml_loop(self):
        import pynvml as nv
        h = nv.nvmlDeviceGetHandleByIndex(self.index)
        smax = nv.nvmlDeviceGetMaxClockInfo(h, nv.NVML_CLOCK_SM)
        get_reasons = getattr(nv, "nvmlDeviceGetCurrentClocksEventReasons", None) or nv.nvmlDeviceGetCurrentClocksThrottleReasons
        while not self.stop_flag.is_set():
            mask = int(get_reasons(h))
            # NVML bit masks of nvmlClocksEventReasons (nvml.h): hw_slowdown, hw_thermal, sw_thermal, sw_power_cap
            flags = ["Active" if mask & bit else "Not Active" for bit in (0x8, 0x40, 0x20, 0x4)]
            self.samples.append(",".join([str(nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM)), str(smax)] + flags))
            time.sleep(0.005)

    def start(self):
        # the timed region can be ~0.1 s: NVML is polled every 5 ms from a thread (the same counters nvidia-smi prints);
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
    """One K-block through the oracle (CPU port of the reference algorithm): every calculator of the workload."""
    dK, nkfft = args
    from oracle import wb_oracle as orc
    W = _oracle_block.W
    data = orc.OracleDataK(_oracle_block.sys, np.asarray(dK, dtype=float), list(nkfft))
    for name, Ef, kw in W.static.values():
        orc.CALCULATORS[name](data, Ef, **kw)
    if W.kubo is not None:
        name, Ef, om, kw = W.kubo
        getattr(orc, name)(data, Ef, omega=om, **kw)
    return data.nk


def _oracle_init(key):
    for v in ("OMP_NUM_THREADS", "OPENBLAS_NUM_THREADS", "MKL_NUM_THREADS"):
        os.environ[v] = "1"
    _oracle_block.W = workloads(key)
    _oracle_block.sys = oracle_system(_oracle_block.W.system)


def k_list(W):
    """(Kp_fullBZ, factor) of the workload's K-list: host-side grid logic only"""
    import wannierberri_b200 as wb

    class _Lat:   # Grid reads the lattice only; NKdiv x NKFFT are given, nothing is derived from the system
        real_lattice = np.eye(3)
        pointgroup = None
        periodic = (True, True, True)
    return wb.Grid(_Lat(), NKdiv=W.NKdiv, NKFFT=W.NKFFT).K_arrays()


def cpu_baseline_1core(W):
    """oracle on 1 core, one K-block of the same workload (same system, calculators and energy axes); for the heavy
    configurations the block is a smaller FFT sub-grid at the same K-list shift (stated in `sample`)."""
    _oracle_init(W.key)
    shifts, _ = k_list(W)
    dK = shifts[min(1, len(shifts) - 1)] * (np.array(W.NKFFT) / np.array(W.cpu_nkfft))   # the same k-space offset of the sub-grid
    t0 = time.perf_counter()
    nk = _oracle_block((dK, W.cpu_nkfft))
    dt = time.perf_counter() - t0
    reduced = "" if W.cpu_nkfft == W.NKFFT else f" (bounded sample: the GPU arm's blocks are NKFFT={W.NKFFT}; cost per k-point is the same)"
    return {"value": nk / dt, "unit": UNIT, "cores": 1, "kind": "port",
            "sample": f"1 K-block of NKFFT={W.cpu_nkfft} ({nk} k-points) of the same workload{reduced}, {dt:.1f} s; "
                      "oracle/wb_oracle.py (numpy restatement of the reference, per-k Python loops as in the reference)"}


def run_reference_arm(args):
    """`--impl reference`: the CPU port on all host cores, on the SAME K-blocks the GPU arm evaluates (for the headline:
    NKFFT = 20^3 sub-grids at the shifts of the 400^3 grid's K-list, same calculators and Fermi levels); each step = one
    K-block per core (the reference parallelises over K-blocks, run_grid.py:258-265)."""
    import multiprocessing as mp
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    W = workloads(args.config)
    cores = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
    shifts, factors = k_list(W)
    nkfft = W.NKFFT if args.config in ("headline", "1") else W.cpu_nkfft
    scale = np.array(W.NKFFT) / np.array(nkfft)
    cursor = [0]
    with mp.Pool(cores, initializer=_oracle_init, initargs=(args.config,)) as pool:
        def step():
            idx = (cursor[0] + np.arange(cores)) % len(factors)   # the GPU arm's K-list, consecutive K-blocks
            cursor[0] += cores
            return sum(pool.map(_oracle_block, [(shifts[i] * scale, nkfft) for i in idx], chunksize=1))
        for _ in range(args.warmup):
            step()
        t0 = time.perf_counter()
        nk = 0
        for _ in range(args.steps):
            nk += step()
        dt = time.perf_counter() - t0
    value = nk / dt
    sample = (f"{cores} K-blocks of NKFFT={nkfft} ({int(np.prod(nkfft))} k-points each) per step, consecutive entries of the "
              f"NKdiv={W.NKdiv} K-list (the GPU arm's list), one per core: multiprocessing.Pool({cores}), BLAS threads = 1")
    cfg = config_dict(W, args.blocks or W.blocks, args.gpus)
    cfg["reference_arm_step"] = {"blocks_per_step": cores, "NKFFT": nkfft, "kpoints_per_step": cores * int(np.prod(nkfft))}
    print(json.dumps({
        "impl": "reference", "metric": W.metric(), "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": data_string(W), "config": cfg,
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0}))


def data_string(W):
    return {"fe": "tests-data system (Fe_W90 of the reference, 18 WF, nR=95)", "te": "tests-data system (Te_qe of the reference, 24 WF, nR=65)",
            "synth32": "synthetic seeded random model (SURVEY 8(d) config 4: 32 WF, nR=125, Ham + AA)",
            "synth128": "synthetic seeded random model (SURVEY 8(d) config 5: 128 WF, 3959 R-vectors, Ham only)"}[W.system] + \
        f"; K-block shifts of the NKdiv={W.NKdiv} grid"


# ------------------------------------------------------------------------------------------ GPU arm
def init_dist(dev):
    import torch
    import torch.distributed as dist
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


def fp64_peaks(local):
    import ctypes as C
    from wannierberri_b200 import _lib
    fp64 = {}
    for kind, name in ((0, "dfma"), (1, "dmma")):
        t = C.c_double()
        _lib.check(_lib.lib().wbgpu_fp64_peak(local, kind, C.byref(t)))
        fp64[name] = t.value
    return fp64


def run_strong(args):
    """`--scaling strong`: the whole BASELINE config-2 grid (8000 K-blocks of 20^3 = 6.4e7 k-points; AHC + Morb, 2000 Fermi
    levels) through run(), total work fixed, K-list sharded over the ranks inside run(); a step = one run()."""
    import torch
    import torch.distributed as dist
    import wannierberri_b200 as wb
    world, rank, local = (int(os.environ.get(k, "0")) for k in ("WORLD_SIZE", "RANK", "LOCAL_RANK"))
    world = max(world, 1)
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        init_dist(dev)
    W = workloads("2" if args.config == "headline" else args.config)
    system = build_system(W.system)
    st = wb.calculators.static
    calcs = {k: getattr(st, name)(Efermi=Ef, **kw) for k, (name, Ef, kw) in W.static.items()}
    grid = wb.Grid(system, NKdiv=W.NKdiv, NKFFT=W.NKFFT)
    nk_total = int(np.prod(W.NKdiv)) * W.nk_block
    kw = dict(use_irred_kpt=False, symmetrize=False, write_files=False, device=local)

    def step():
        return wb.run(system, grid, calcs, **kw)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    wb.run(system, wb.Grid(system, NKdiv=[2, 2, 2], NKFFT=W.NKFFT), calcs, **kw)   # plan + buffers
    for _ in range(max(args.warmup, 1)):
        step()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    eng = wb.data_K.engine_for(system, local)
    l0 = eng.kernel_launches
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record()
    for _ in range(args.steps):
        res = step()
    e1.record()
    barrier()
    ms = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    ms = float(ms.item())
    launches = eng.kernel_launches - l0
    clocks = sampler.stop() if rank == 0 else None
    value = nk_total * args.steps / (ms * 1e-3)
    nout = sum(int(np.prod(r.data.shape)) for r in res.results.values())
    if rank == 0:
        shifts, factors = grid.K_arrays()
        print(json.dumps({
            "metric": W.metric(), "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 1),
            "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64",
            "data": data_string(W), "config": config_dict(W, int(np.prod(W.NKdiv)), world, "strong"),
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": int(shifts.nbytes + factors.nbytes), "d2h_bytes_per_step": nout * 8,
                    "api": "wannierberri_b200.run() (host K-list in, EnergyResult out; the timed region IS the public call)"},
            "gpu_launches": int(launches), "clocks": clocks,
            "check": {k: float(np.abs(r.data).max()) for k, r in res.results.items()}}))
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200")
    ap.add_argument("--config", default="headline", help="headline (default) or a BASELINE configuration 1..5")
    ap.add_argument("--scaling", default="weak", choices=["weak", "strong"])
    ap.add_argument("--blocks", type=int, default=0, help="K-blocks per GPU per step (0 = the configuration's default)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference_arm(args)
    if args.scaling == "strong":
        return run_strong(args)

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
        init_dist(dev)

    W = workloads(args.config)
    system = build_system(W.system)
    nw = system.num_wann
    st, dyn = wb.calculators.static, wb.calculators.dynamic
    calcs = {k: getattr(st, name)(Efermi=Ef, **kw) for k, (name, Ef, kw) in W.static.items()}
    specs = [s for c in calcs.values() for s in c.specs()]
    kcalc = getattr(dyn, W.kubo[0])(Efermi=W.kubo[1], omega=W.kubo[2], **W.kubo[3]) if W.kubo else None
    kspec = kcalc.spec() if kcalc else None
    eng = wb.Engine(system, device=local)
    flags = [s.formula for s in specs] + ([kspec.formula_flag] if kspec else [])
    eng.plan(W.NKFFT, flags, external_terms=True)

    # this rank's K-blocks: a contiguous chunk of the grid's K-list (shifts of the FFT sub-grid)
    shifts, factors = k_list(W)
    nb = args.blocks or W.blocks
    lo = (rank * nb) % len(factors)
    idx = (lo + np.arange(nb)) % len(factors)
    dK_h = np.ascontiguousarray(shifts[idx])
    w_h = np.ascontiguousarray(factors[idx])
    dK_d, w_d = torch.from_numpy(dK_h).pin_memory().to(dev), torch.from_numpy(w_h).pin_memory().to(dev)
    nstat = sum(s.size for s in specs)
    nout = nstat + (int(_lib.lib().wbgpu_kubo_size(C.byref(kspec))) if kspec else 0)
    out_d = torch.zeros(nout, dtype=torch.float64, device=dev)
    kpts_step = nb * W.nk_block * world

    def step_dev():
        if specs:
            eng.scan_dev(dK_d, w_d, specs, out_d)
        if kspec:
            eng.kubo_scan_dev(dK_d, w_d, kspec, kcalc.Efermi, kcalc.omega, out_d[nstat:])
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

    if kspec:   # contributing band pairs per k-point of the Kubo scan, from the eigenvalues of this rank's first K-block
        Eprobe = eng.eig(dK_h[0])
        lo_e, hi_e = float(np.min(kcalc.Efermi)), float(np.max(kcalc.Efermi))
        occ = (Eprobe <= hi_e)[:, :, None] & (Eprobe > lo_e)[:, None, :] & (np.arange(nw)[:, None] < np.arange(nw)[None, :])[None]
        W.pairs_per_k = float(occ.sum(axis=(1, 2)).mean())
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
        arrays = eng.scan(dK_h, w_h, specs) if specs else []
        if kspec:
            arrays = arrays + [np.ascontiguousarray(eng.kubo_scan(dK_h, w_h, kspec, kcalc.Efermi, kcalc.omega)).view(np.float64)]
        if world > 1:
            t = torch.from_numpy(np.concatenate([a.ravel() for a in arrays])).to(dev)
            dist.all_reduce(t)
            arrays = [t.cpu().numpy()]
        host_out["last"] = arrays

    step_e2e()
    ms_e2e = timed(step_e2e, args.steps)
    api = "wannierberri_b200.Engine.scan -> wbgpu_static_scan (host pointers)" if specs else ""
    if kspec:
        api = (api + " + " if api else "") + "Engine.kubo_scan -> wbgpu_kubo_scan (host pointers)"
    e2e = {"value": kpts_step * args.steps / (ms_e2e * 1e-3), "unit": UNIT,
           "h2d_bytes_per_step": int(dK_h.nbytes + w_h.nbytes) * (bool(specs) + bool(kspec)), "d2h_bytes_per_step": int(nout * 8), "api": api}

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
    stage_calls = {STAGES[i]: max(int(calls[i]) // args.steps, 1) for i in range(5)}
    flops = W.flops_per_k(nw)
    cands = [s for s in ("fourier", "eigh", "rotate", "scan") if flops.get(s, 0) > 0]
    dominant = max(cands, key=lambda s: stage_ms[s])

    if rank == 0:
        fp64 = fp64_peaks(local)
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except OSError:
            pass
        hbm_peak = peaks.get("hbm_gbs", 6650.0)
        kpts_rank = nb * W.nk_block
        ninst = stage_calls[dominant]                      # timed instances of the stage per step (sub-batches)
        k_per_launch = kpts_rank / ninst
        dur = stage_ms[dominant] / ninst * 1e-3
        achieved = flops[dominant] * k_per_launch / dur / 1e12
        peak64 = max(fp64.values())
        pipe = {"rotate": "FP64 tensor (mma.sync.m8n8k4.f64)", "eigh": "FP64 (DFMA; DMMA back-transformation for nw > 32)",
                "fourier": "FP64 (DFMA)", "scan": "FP64 (DFMA)"}[dominant]
        note = ("stage = the kernels between two CUDA events on the library's stream; algorithmic flops per k-point of "
                "SURVEY.md section 8(d)")
        if dominant == "scan":
            note += (f" (scan: difference-form Kubo accumulation, {W.pairs_per_k:.1f} contributing band pairs per k-point probed "
                     f"from one K-block; the reference's dense contraction would be {W.scan_dense_flops(nw):.3g} flops per k-point)")
        roofline = {"kernel": dominant, "bound": "tensor", "pipe": pipe, "achieved": achieved, "peak": peak64, "unit": "TFLOP/s",
                    "frac": achieved / peak64, "traffic": None,
                    "peak_source": f"measured in this run: DFMA {fp64['dfma']:.1f}, DMMA(m8n8k4) {fp64['dmma']:.1f} TFLOP/s "
                                   "(MEASURED_PEAKS.json holds no FP64 figure; tcgen05 has no FP64 kind)",
                    "algorithmic_flops_per_k": flops[dominant], "avg_launch_ms": dur * 1e3, "kpoints_per_launch": k_per_launch,
                    "note": note}
        step_s = sum(stage_ms.values()) * 1e-3
        bpk = W.bytes_per_k(nw)
        roofline_hbm = {"bound": "hbm", "achieved": bpk * kpts_rank / step_s / 1e9, "peak": hbm_peak, "unit": "GB/s",
                        "frac": bpk * kpts_rank / step_s / 1e9 / hbm_peak,
                        "peak_source": "MEASURED_PEAKS.json" if peaks else "fallback (B200_PROFILING.md)",
                        "note": f"whole pipeline, algorithmic X(k) bytes (16*nw^2*{W.nmat} per k-point) / sum of stage times"}
        total_flops = sum(flops.values())
        roofline_all = {"bound": "fp64", "achieved": total_flops * kpts_rank / step_s / 1e12, "peak": peak64, "unit": "TFLOP/s",
                        "frac": total_flops * kpts_rank / step_s / 1e12 / peak64,
                        "note": "whole pipeline, sum of the algorithmic flops of all stages / sum of stage times"}
        cpu = None if (args.no_cpu_baseline or world > 1) else cpu_baseline_1core(W)   # rank 0 at N = 1 only
        print(json.dumps({
            "metric": W.metric(), "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f64", "data": data_string(W),
            "config": config_dict(W, nb, world), "e2e": e2e, "gpu_launches": int(launches), "clocks": clocks,
            "roofline": roofline, "roofline_hbm": roofline_hbm, "roofline_pipeline": roofline_all, "stage_ms_per_step": stage_ms,
            "stage_instances_per_step": stage_calls, "fp64_peak_tflops": fp64, "cpu_baseline": cpu}))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
