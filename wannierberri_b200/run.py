"""`run()` with the signature of the reference (run_grid.py:118-142).  The K-block loop
(`process`, run_grid.py:32-115) and its Ray fan-out are replaced by: shard the K-block list over the
ranks of `torch.distributed` (one process per GPU), evaluate each shard with ONE call into
libwbgpu.so, combine with ONE all-reduce of the Fermi-scan arrays."""
import glob
import os
import pickle
import shutil

import numpy as np

from .calculators.static import adapt as adapt_static
from .calculators import dynamic as _dyn
from . import _lib
from .data_K import engine_for, check_parameters_K, Data_K_R
from .result import ResultDict
from .system import as_system


def _dist():
    try:
        import torch.distributed as dist
        if dist.is_available() and dist.is_initialized():
            return dist
    except ImportError:
        pass
    return None


def shard_bounds(n, rank, world):
    """Contiguous chunk [lo, hi) of n K-blocks for `rank` of `world`."""
    base, rem = divmod(n, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def k_list_arrays(grid, use_irred_kpt):
    """(Kp_fullBZ[nK,3], factor[nK]) from this package's Grid or from the reference's."""
    if hasattr(grid, "K_arrays"):
        return grid.K_arrays(use_symmetry=use_irred_kpt)
    K_list = grid.get_K_list(use_symmetry=use_irred_kpt)
    return (np.array([K.Kp_fullBZ for K in K_list], dtype=float).reshape(-1, 3),
            np.array([K.factor for K in K_list], dtype=float))


def _pointgroup_of(system):
    """`system.pointgroup`; a system on which `set_pointgroup` was never called has the identity group (the reference's
    `PointGroup()` with no generators, system/system.py:140-170), so `use_irred_kpt` / `symmetrize` are no-ops on it."""
    pg = getattr(system, "pointgroup", None)
    if pg is None:
        from .symmetry import PointGroup
        pg = PointGroup((), real_lattice=system.real_lattice)
    return pg


def run(system, grid, calculators, adpt_num_iter=0, use_irred_kpt=True, symmetrize=True, fout_name="result",
        suffix="", parameters_K=None, file_Klist_path=None, restart=False, allow_restart=False, dump_results=False,
        restart_iteration=-1, Klist_part=10, parallel=True, print_Kpoints=False, adpt_mesh=2, adpt_fac=1,
        print_progress_step_time=5, print_progress_step_percent=1, data_k_class=None, k_batch=50,
        device=None, write_files=True):
    """Integrate `calculators` over the k-grid.  Returns a `ResultDict` of `EnergyResult`.

    Signature and defaults of the reference (run_grid.py:118-142): symmetry-irreducible K-points and symmetrised
    results by default (`use_irred_kpt=True` forces `symmetrize=True`, run_grid.py:233-234); after every iteration
    each quantity is written to `<fout_name>-<key>[-<suffix>]_iter-NNNN.npz` / `.dat` by rank 0 (run_grid.py:368;
    `write_files=False`, an extension of this package, switches that off).

    `dump_results=True` (run_grid.py:62-63, 242-243): per-K-point results are pickled to
    `<file_Klist_path>/_Kp-<ik>.pickle` and dropped from memory; implies `allow_restart`.

    `parameters_K`: `fftlib` is accepted (and has no effect: the R->k transform is the CUDA one); non-default
    `Emin` / `Emax` / `random_gauge` raise, as does a `data_k_class` other than this package's (no CPU fallback)."""
    from .data_K import DataKHost
    if data_k_class is not None and not (isinstance(data_k_class, type) and issubclass(data_k_class, DataKHost)):
        raise NotImplementedError(f"data_k_class {getattr(data_k_class, '__name__', data_k_class)}: only this package's "
                                  "Data_K_R (or a subclass of its DataKHost interface, for plug-in calculators) runs here")
    if data_k_class is None:
        data_k_class = Data_K_R
    if use_irred_kpt:   # run_grid.py:233-234: the irreducible wedge alone is meaningless without symmetrisation
        symmetrize = True
    if dump_results:
        allow_restart = True
    if adpt_num_iter != 0 or restart or allow_restart:   # per-K-point results are kept: the refinement loop
        return _run_adaptive(system, grid, calculators, adpt_num_iter, adpt_mesh, adpt_fac, fout_name, suffix, parallel,
                             device, write_files, symmetrize, use_irred_kpt, parameters_K,
                             dict(restart=restart, allow_restart=allow_restart, restart_iteration=restart_iteration,
                                  Klist_part=Klist_part, file_Klist_path=file_Klist_path, dump_results=dump_results))
    check_parameters_K(parameters_K)
    system = as_system(system)
    pointgroup = _pointgroup_of(system)
    calcs, dyn_calcs, plug_calcs = {}, {}, {}
    for key, c in calculators.items():
        # static.SHC and dynamic.SHC share their class name: a Kubo calculator is the one that carries a frequency axis
        dynamic = isinstance(c, _dyn.DynamicCalculator) or (type(c).__name__ in _dyn._BY_NAME and hasattr(c, "omega")
                                                             and not hasattr(c, "fder"))
        if not dynamic:
            try:
                c = adapt_static(c)
            except KeyError:   # not one of the scans of the CUDA kernels: a user's calculator, called per K-block
                if not callable(c):
                    raise ValueError(f"calculator {key} ({type(c).__name__}) is neither available on the GPU path nor callable")
            if not getattr(c, "allow_grid", True):
                raise ValueError(f"Calculator {key} is not compatible with a grid")
            if getattr(c, "is_plugin", True):
                plug_calcs[key] = c
            else:
                calcs[key] = c
            continue
        c = _dyn.adapt(c)
        if not c.allow_grid:
            raise ValueError(f"Calculator {key} is not compatible with a grid")
        dyn_calcs[key] = c

    dist = _dist() if parallel else None
    rank, world = (dist.get_rank(), dist.get_world_size()) if dist else (0, 1)
    if device is None:
        import torch
        device = torch.cuda.current_device() if torch.cuda.is_available() else 0

    shifts, factors = k_list_arrays(grid, use_irred_kpt)
    lo, hi = shard_bounds(len(factors), rank, world)

    specs, owner, tspecs, towner = [], [], [], []
    for key, c in calcs.items():
        for s in c.specs():
            if getattr(system, "force_internal_terms_only", False):
                s.external_terms = 0
            (tspecs if c.tetra else specs).append(s)
            (towner if c.tetra else owner).append(key)
    internal_only = getattr(system, "force_internal_terms_only", False)
    kspecs = {}
    for key, c in dyn_calcs.items():
        ks = c.spec()
        if internal_only:
            ks.external_terms = 0
        kspecs[key] = ks
    external = any(s.external_terms for s in specs + tspecs) or any(ks.external_terms for ks in kspecs.values())
    formulae = {int(s.formula) for s in specs + tspecs} | {ks.formula_flag for ks in kspecs.values()} | {_lib.IDENTITY}
    if (calcs or dyn_calcs) and data_k_class is not Data_K_R:
        raise NotImplementedError("the library's calculators run on this package's Data_K_R only (no CPU fallback)")
    engine = engine_for(system, device) if (calcs or dyn_calcs) else None
    if engine is not None:
        engine.plan(np.array(grid.FFT, dtype=int), formulae, external_terms=external)
    arrays = engine.scan(shifts[lo:hi], factors[lo:hi], specs) if specs else []
    if tspecs:  # tetrahedron method: the cell around every k-point is KpointBZparallel.dK_fullBZ
        dK_cell = 1. / (np.array(grid.div, dtype=float) * np.array(grid.FFT, dtype=float))
        arrays = arrays + engine.scan_tetra(shifts[lo:hi], factors[lo:hi], dK_cell, tspecs)
        owner = owner + towner
    # Kubo scans: one call each (their accumulators are large; they do not share the event pass of the static scans)
    karrays = [engine.kubo_scan(shifts[lo:hi], factors[lo:hi], ks, dyn_calcs[key].Efermi, dyn_calcs[key].omega)
               for key, ks in kspecs.items()]
    # plug-in calculators (user formulae / calculators: SURVEY.md section 8(b), hooks #1 and #2): called K-block by
    # K-block on a GPU-resident Data_K_R as the reference's `process` does (run_grid.py:32-115, 258-265)
    plug_res = {key: None for key in plug_calcs}
    for i in range(lo, hi):
        if not plug_calcs:
            break
        data_K = data_k_class(system, shifts[i], grid, device=device)
        for key, c in plug_calcs.items():
            r = c(data_K) * factors[i]
            plug_res[key] = r if plug_res[key] is None else plug_res[key] + r
    if plug_calcs and any(r is None for r in plug_res.values()):   # a rank without K-blocks: the shape comes from K-block 0
        data_K = data_k_class(system, shifts[0], grid, device=device)
        for key, c in plug_calcs.items():
            if plug_res[key] is None:
                plug_res[key] = c(data_K) * 0.
    nstatic = len(arrays)
    arrays = arrays + [np.ascontiguousarray(a).view(np.float64) for a in karrays]
    nkubo = len(karrays)
    arrays = arrays + [np.ascontiguousarray(plug_res[key].data, dtype=np.float64) for key in plug_calcs]

    if dist and world > 1:
        import torch
        flat = np.concatenate([a.ravel() for a in arrays])
        dev = torch.device("cuda", device) if dist.get_backend() == "nccl" else torch.device("cpu")
        t = torch.from_numpy(flat).to(dev)
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        flat = t.cpu().numpy()
        off = 0
        for i, a in enumerate(arrays):
            arrays[i] = flat[off:off + a.size].reshape(a.shape)
            off += a.size

    results = {}
    for key, c in calcs.items():
        mine = [a for a, o in zip(arrays[:nstatic], owner) if o == key]
        results[key] = c.result(mine, system.cell_volume)
    for (key, c), a, raw in zip(dyn_calcs.items(), arrays[nstatic:nstatic + nkubo], karrays):
        results[key] = c.result(a.view(raw.dtype).reshape(raw.shape))
    for key, a in zip(plug_calcs, arrays[nstatic + nkubo:]):
        results[key] = _as_energy_result(plug_res[key], a)
    if symmetrize and pointgroup.size > 1:  # linear: applied once to the weighted sum instead of per K-point (run_grid.py:258-265)
        results = {key: r.symmetrized(pointgroup) for key, r in results.items()}
    res = ResultDict(results)
    if write_files and rank == 0:
        res.savedata(prefix=fout_name, suffix=suffix, i_iter=0)
    return res


def _as_energy_result(r, data):
    """the summed result of a plug-in calculator with the all-reduced data, as this package's EnergyResult (a result
    object of the reference is recognised by its attributes)"""
    from .result import EnergyResult
    if isinstance(r, EnergyResult):
        out = r * 1.
        out.data = np.asarray(data).reshape(r.data.shape)
        return out
    for attr in ("Energies", "data", "transformTR", "transformInv"):
        if not hasattr(r, attr):
            raise ValueError(f"the result of a plug-in calculator must be an EnergyResult (got {type(r).__name__})")
    return EnergyResult(list(r.Energies), np.asarray(data).reshape(np.shape(r.data)), transformTR=r.transformTR,
                        transformInv=r.transformInv, rank=getattr(r, "rank", None),
                        E_titles=getattr(r, "E_titles", ("Efermi",)), comment=getattr(r, "comment", "undocumented"),
                        save_mode=getattr(r, "save_mode", "bin+txt"), smoothers=getattr(r, "smoothers", (None,)))


def _write_factors(path, factors, it):   # run_grid.py:420-422
    with open(os.path.join(path, f"factors_iter-{it:08d}.npy"), "wb") as f:
        np.save(f, factors)


def _read_factors(path, it):   # run_grid.py:425-440
    if it >= 0:
        return it, np.load(os.path.join(path, f"factors_iter-{it:08d}.npy"))
    have = sorted(int(f.split("-")[-1].split(".")[0]) for f in glob.glob(os.path.join(path, "factors_iter-*.npy")))
    want = max(have[-1] + it + 1, 0)
    return _read_factors(path, max(i for i in have if i <= want) if want > 0 else 0)


def _run_adaptive(system, grid, calculators, adpt_num_iter, adpt_mesh, adpt_fac, fout_name, suffix, parallel, device,
                  write_files, symmetrize, use_irred_kpt, parameters_K, restart_opts=None):
    """The refinement loop of the reference (run_grid.py:303-387) on per-K-block results from the GPU
    (`wbgpu_static_scan_blocks`): evaluate the new K-points, update the weighted sum, pick the `adpt_fac` points with
    the largest contribution by every criterion of `ResultDict.max`, divide them `adpt_mesh`-fold, repeat."""
    from .grid import KpointBZparallel, exclude_equiv_points
    ro = dict(restart=False, allow_restart=False, restart_iteration=-1, Klist_part=10, file_Klist_path=None,
              dump_results=False)
    ro.update(restart_opts or {})
    Klist_dir = ro["file_Klist_path"] if ro["file_Klist_path"] is not None else "_tmp_wb"   # run_grid.py:244-246
    file_Klist = os.path.join(Klist_dir, "K_list.pickle")
    check_parameters_K(parameters_K)
    system = as_system(system)
    pointgroup = _pointgroup_of(system)
    periodic = getattr(system, "periodic", (True, True, True))
    calcs, slow_calcs = {}, {}   # slow: tetrahedron / Kubo calculators (their own per-K-block entry points)
    for key, c in calculators.items():
        if isinstance(c, _dyn.DynamicCalculator) or (type(c).__name__ in _dyn._BY_NAME and hasattr(c, "omega")
                                                       and not hasattr(c, "fder")):
            slow_calcs[key] = _dyn.adapt(c)
            continue
        c = adapt_static(c)
        (slow_calcs if c.tetra else calcs)[key] = c
    dist = _dist() if parallel else None
    rank, world = (dist.get_rank(), dist.get_world_size()) if dist else (0, 1)
    if device is None:
        import torch
        device = torch.cuda.current_device() if torch.cuda.is_available() else 0

    NKFFT = np.array(grid.FFT, dtype=int)
    shifts, factors = k_list_arrays(grid, use_irred_kpt)
    dK0 = 1. / np.array(grid.div, dtype=float)
    K_list = [KpointBZparallel(s * NKFFT, dK0, NKFFT, f, pointgroup=pointgroup if use_irred_kpt else None)
              for s, f in zip(shifts, factors)]
    start_iter, nk_saved = 0, 0
    restored = None
    if ro["restart"]:   # run_grid.py:271-288: the evaluated K-points with their results, the factors of the chosen iteration
        K_list = []
        with open(file_Klist, "rb") as fr:
            while True:
                try:
                    K_list += pickle.load(fr)
                except EOFError:
                    break
        nk_saved = len(K_list)
        start_iter, fac = _read_factors(Klist_dir, ro["restart_iteration"])
        fac = np.hstack([fac, np.zeros(len(K_list) - len(fac))])
        for K, f in zip(K_list, fac):
            K.factor = float(f)
        restored = fac
    if adpt_num_iter < 0:  # run_grid.py:303-304
        adpt_num_iter = -adpt_num_iter * np.prod(grid.div) / np.prod(adpt_mesh) / adpt_fac / 3
    adpt_num_iter = int(round(adpt_num_iter))
    adpt_mesh = np.array([adpt_mesh] * 3 if np.ndim(adpt_mesh) == 0 else adpt_mesh, dtype=int)
    if np.max(adpt_mesh) <= 1:
        adpt_num_iter = 0

    specs, owner = [], []
    for key, c in calcs.items():
        for s in c.specs():
            if getattr(system, "force_internal_terms_only", False):
                s.external_terms = 0
            specs.append(s)
            owner.append(key)
    internal_only = getattr(system, "force_internal_terms_only", False)
    slow_specs = {}
    for key, c in slow_calcs.items():
        sp = [c.spec()] if isinstance(c, _dyn.DynamicCalculator) else c.specs()
        for s in sp:
            if internal_only:
                s.external_terms = 0
        slow_specs[key] = sp
    flags = {int(s.formula) for s in specs} | {_lib.IDENTITY}
    for key, c in slow_calcs.items():
        flags |= {s.formula_flag if isinstance(c, _dyn.DynamicCalculator) else int(s.formula) for s in slow_specs[key]}
    external = any(s.external_terms for s in specs) or any(s.external_terms for sp in slow_specs.values() for s in sp)
    engine = engine_for(system, device)
    engine.plan(NKFFT, flags, external_terms=external)

    def eval_slow_batch(key, Ks):
        """all new K-points of this rank through the tetrahedron / Kubo scans in ONE call (per-K-block histograms /
        accumulators in the library: wbgpu_static_scan_tetra_blocks, wbgpu_kubo_scan_blocks)"""
        if not Ks:
            return []
        c, sp = slow_calcs[key], slow_specs[key]
        dK = np.array([K.Kp_fullBZ for K in Ks], dtype=float).reshape(-1, 3)
        if isinstance(c, _dyn.DynamicCalculator):
            return [[np.ascontiguousarray(a)] for a in engine.kubo_scan_blocks(dK, sp[0], c.Efermi, c.omega)]
        cells = np.array([K.dK_fullBZ for K in Ks], dtype=float).reshape(-1, 3)
        per_spec = engine.scan_tetra_blocks(dK, cells, sp)
        return [[a[j] for a in per_spec] for j in range(len(Ks))]

    result_all = None
    factors_old = None
    if restored is not None:
        for K in K_list:
            contrib = K.get_result() * K.factor
            result_all = contrib if result_all is None else result_all + contrib
        factors_old = restored
    elif ro["allow_restart"] and rank == 0:   # run_grid.py:295-298
        shutil.rmtree(Klist_dir, ignore_errors=True)
        os.makedirs(Klist_dir)
        _write_factors(Klist_dir, np.array([K.factor for K in K_list]), 0)
    for i_iter in range(adpt_num_iter + 1):
        new = [i for i, K in enumerate(K_list) if not K.was_evaluated_flag]
        # the new K-points are sharded over the ranks; every rank then holds all per-K-point results
        lo, hi = shard_bounds(len(new), rank, world)
        dK_new = np.array([K_list[i].Kp_fullBZ for i in new], dtype=float).reshape(-1, 3)
        mine = engine.scan_blocks(dK_new[lo:hi], specs) if (hi > lo and specs) else [np.zeros((0,) + s.shape) for s in specs]
        slow = {}   # key -> list over the new K-points of this rank's shard of [array per spec]
        for key in slow_calcs:
            slow[key] = eval_slow_batch(key, [K_list[i] for i in new[lo:hi]])
        if dist and world > 1 and slow_calcs:
            import torch
            dev = torch.device("cuda", device) if dist.get_backend() == "nccl" else torch.device("cpu")
            for key, c in slow_calcs.items():
                shapes = [s.shape for s in slow_specs[key]]
                cplx = isinstance(c, _dyn.DynamicCalculator) and slow_specs[key][0].is_complex
                gathered = []
                for isp, shp in enumerate(shapes):
                    buf = np.zeros((len(new),) + shp, dtype=complex if cplx else float)
                    for j in range(hi - lo):
                        buf[lo + j] = slow[key][j][isp]
                    t = torch.from_numpy(np.ascontiguousarray(buf).view(np.float64)).to(dev)
                    dist.all_reduce(t, op=dist.ReduceOp.SUM)
                    gathered.append(t.cpu().numpy().view(buf.dtype).reshape(buf.shape))
                slow[key] = [[gathered[isp][j] for isp in range(len(shapes))] for j in range(len(new))]
            slow_lo = 0
        else:
            slow_lo = lo
        if dist and world > 1:
            import torch
            dev = torch.device("cuda", device) if dist.get_backend() == "nccl" else torch.device("cpu")
            full = []
            for a, s in zip(mine, specs):
                buf = np.zeros((len(new),) + s.shape)
                buf[lo:hi] = a
                t = torch.from_numpy(buf).to(dev)
                dist.all_reduce(t, op=dist.ReduceOp.SUM)   # disjoint slices: a gather written as a sum
                full.append(t.cpu().numpy())
            mine = full
        result_sum_iter = None
        for j, i in enumerate(new):
            res = {key: c.result([a[j] for a, o in zip(mine, owner) if o == key], system.cell_volume)
                   for key, c in calcs.items()}
            for key, c in slow_calcs.items():
                arrs = slow[key][j - slow_lo]
                res[key] = c.result(arrs[0]) if isinstance(c, _dyn.DynamicCalculator) else c.result(arrs, system.cell_volume)
            if symmetrize and pointgroup.size > 1:   # per K-point here: K.max must see the symmetrised result (run_grid.py:258-265)
                res = {key: r.symmetrized(pointgroup) for key, r in res.items()}
            res = ResultDict(res)
            K_list[i].set_result(res)
            contrib = res * K_list[i].factor
            if ro["dump_results"]:   # run_grid.py:62-63, 321, 416-417; every rank holds all results, rank 0 writes
                K_list[i].set_storage_path(os.path.join(Klist_dir, f"_Kp-{i}.pickle"))
                if rank == 0:
                    K_list[i].dump_result()
                else:
                    K_list[i].result, K_list[i].res_dumped_flag = None, True
            result_sum_iter = contrib if result_sum_iter is None else result_sum_iter + contrib
        fac_now = np.array([K.factor for K in K_list])
        if ro["allow_restart"] and rank == 0:   # run_grid.py:343-348: append the K-points evaluated in this iteration
            with open(file_Klist, "ab") as fw:
                for ink in range(nk_saved, len(K_list), int(ro["Klist_part"])):
                    pickle.dump(K_list[ink:ink + int(ro["Klist_part"])], fw)
            if result_all is not None:
                _write_factors(Klist_dir, fac_now, i_iter + start_iter)
        nk_saved = len(K_list)
        if result_all is None:
            result_all = result_sum_iter
        else:  # run_grid.py:352-360: the points divided in the previous iteration lost their weight
            diff = fac_now[:len(factors_old)] - factors_old
            if result_sum_iter is not None:
                result_all = result_all + result_sum_iter
            for i, d in enumerate(diff):
                if abs(d) > 1.e-8:
                    result_all = result_all + K_list[i].get_result() * d
        factors_old = fac_now
        if write_files and rank == 0 and not (ro["restart"] and i_iter == 0):
            result_all.savedata(prefix=fout_name, suffix=suffix, i_iter=i_iter + start_iter)
        if i_iter >= adpt_num_iter:
            break
        Kmax = np.array([K.max for K in K_list]).T
        select_points = set().union(*(np.argsort(Km)[-adpt_fac:] for Km in Kmax))
        l1 = len(K_list)
        for iK in select_points:
            K_list += K_list[iK].divide(adpt_mesh, periodic=periodic, use_symmetry=use_irred_kpt)
        if use_irred_kpt:
            exclude_equiv_points(K_list, new_points=len(K_list) - l1)
    return result_all
