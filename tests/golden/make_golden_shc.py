#!/usr/bin/env python
"""Golden fixture of the Kubo spin Hall conductivity (dynamic.SHC, SHC_type = ryoo / qiao / simple) from the UNMODIFIED
upstream reference on its `random` test system (which carries SA, SHA, SR, SH, SHR); asserts that the live run
reproduces the reference's own golden files random-opt_SHC{qiao,ryoo}_iter-0000.npz (tests/test_run.py:653-669).
The spin matrices are added to tests/golden/random_system.npz.

    cd /tmp && PYTHONPATH=/root/reference:/root/repo/oracle/stubs python /root/repo/tests/golden/make_golden_shc.py
"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from make_golden import REF, OUT, run_ref, dump_system, System_R  # noqa: E402
from wannierberri.calculators import dynamic as dyn  # noqa: E402


def main():
    system = System_R.from_npz(path=os.path.join(REF, "tests", "data", "random"), legacy=True)
    dump_system(system, "random_system.npz", ("Ham", "AA", "BB", "CC", "SS", "SA", "SHA", "SR", "SH", "SHR"))
    p_ref = dict(Efermi=np.array([17.0, 18.0]), omega=np.arange(0.0, 7.1, 1.0), smr_fixed_width=0.20, smr_type="Gaussian")
    p_in = dict(Efermi=np.linspace(-2, 2, 9), omega=np.arange(0.0, 7.1, 1.0), smr_fixed_width=0.20, smr_type="Lorentzian")
    calcs = dict(ref_qiao=dyn.SHC(SHC_type="qiao", **p_ref), ref_ryoo=dyn.SHC(SHC_type="ryoo", **p_ref),
                 in_qiao=dyn.SHC(SHC_type="qiao", **p_in), in_ryoo=dyn.SHC(SHC_type="ryoo", **p_in),
                 in_simple=dyn.SHC(SHC_type="simple", **p_in),
                 in_ryoo_thresh=dyn.SHC(SHC_type="ryoo", degen_thresh=0.3, **p_in))
    grid, res = run_ref(system, [6, 6, 6], [3, 3, 3], calcs)
    out = dict(NK=np.array([6, 6, 6]), NKFFT=np.array([3, 3, 3]), ref_Efermi=p_ref["Efermi"], in_Efermi=p_in["Efermi"],
               omega=p_ref["omega"])
    for t in ("qiao", "ryoo"):
        ref = np.load(os.path.join(REF, "tests/reference/integrate_files", f"random-opt_SHC{t}_iter-0000.npz"))["data"]
        got = res.results["ref_" + t].data
        err = np.abs(got - ref).max() / max(np.abs(ref).max(), 1e-300)
        print(f"random-opt_SHC{t}: live reference run vs reference golden file: rel err {err:.2e} (max |ref| {np.abs(ref).max():.3e})")
        assert err < 1e-8 or np.abs(ref).max() < 1e-12
        out["upstream_golden_" + t] = ref
    for q in calcs:
        out[q] = res.results[q].data
    np.savez_compressed(os.path.join(OUT, "golden_random_shc.npz"), **out)
    print("written")


if __name__ == "__main__":
    sys.exit(main())
