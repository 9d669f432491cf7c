#!/usr/bin/env python
"""Fixture of the plug-in test: a USER-DEFINED formula (tests/plugin_formula.py) evaluated by the unmodified reference
(`wannierberri.run` on its own Data_K_R) on the Fe 18-WF system.

    cd /tmp && PYTHONPATH=/root/reference:/root/repo/oracle/stubs:/root/repo/tests python /root/repo/tests/golden/make_golden_plugin.py
"""
import os
import sys

import numpy as np

OUT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, OUT)
sys.path.insert(0, os.path.dirname(OUT))

import wannierberri as wberri  # noqa: E402
from wannierberri.calculators.static import StaticCalculator  # noqa: E402
from make_golden import build_fe  # noqa: E402
from plugin_formula import make_calculators  # noqa: E402

system = build_fe()
Efermi = np.linspace(15.0, 19.0, 9)
grid = wberri.Grid(system, NK=[4, 4, 4], NKFFT=[2, 2, 2])
calcs = make_calculators(StaticCalculator, Efermi)
res = wberri.run(system, grid=grid, calculators=calcs, adpt_num_iter=0, use_irred_kpt=False, symmetrize=False,
                 parallel=False, fout_name=os.path.join("/tmp", "plugin"), print_progress_step_time=1e9)
out = dict(Efermi=Efermi, NK=np.array([4, 4, 4]), NKFFT=np.array([2, 2, 2]))
for key in calcs:
    out[key] = res.results[key].data
    print(key, out[key].shape, np.abs(out[key]).max())
np.savez_compressed(os.path.join(OUT, "golden_plugin.npz"), **out)
