#!/usr/bin/env python
"""Golden fixture of Ohmic_FermiSea (formula InvMass: second comma-derivative of H + generalised derivative) from the
UNMODIFIED upstream reference on the Fe and `random` systems; asserts that the live runs reproduce the reference's own
golden files Fe_W90-conductivity_ohmic_iter-0000.npz and random-conductivity_ohmic_iter-0000.npz.

    cd /tmp && PYTHONPATH=/root/reference:/root/repo/oracle/stubs python /root/repo/tests/golden/make_golden_ohmic_sea.py
"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from make_golden import REF, OUT, build_fe, run_ref, System_R, calc  # noqa: E402


def main():
    out = {}
    fe = build_fe()
    rnd = System_R.from_npz(path=os.path.join(REF, "tests", "data", "random"), legacy=True)
    for tag, system, NK, NKFFT, Ef, fname in (("fe", fe, [4, 4, 4], [2, 2, 2], np.linspace(17, 18, 11), "Fe_W90"),
                                              ("random", rnd, [6, 6, 6], [3, 3, 3], np.linspace(-2, 2, 5), "random")):
        calcs = dict(ohmic=calc.static.Ohmic_FermiSea(Efermi=Ef), ohmic_thresh=calc.static.Ohmic_FermiSea(Efermi=Ef, degen_thresh=0.05),
                     ohmic_tetra=calc.static.Ohmic_FermiSea(Efermi=Ef, tetra=True))
        grid, res = run_ref(system, NK, NKFFT, calcs)
        ref = np.load(os.path.join(REF, "tests/reference/integrate_files", f"{fname}-conductivity_ohmic_iter-0000.npz"))["data"]
        got = res.results["ohmic"].data
        err = np.abs(got - ref).max() / np.abs(ref).max()
        print(f"{fname}-conductivity_ohmic: live reference run vs reference golden file: rel err {err:.2e}")
        assert err < 1e-8
        out[f"{tag}_Efermi"] = Ef
        out[f"{tag}_upstream_golden_ohmic"] = ref
        for q in calcs:
            out[f"{tag}_{q}"] = res.results[q].data
    np.savez_compressed(os.path.join(OUT, "golden_ohmic_sea.npz"), **out)
    print("written")


if __name__ == "__main__":
    sys.exit(main())
