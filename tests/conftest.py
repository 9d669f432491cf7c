import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: test needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def golden_dir():
    return GOLDEN


@pytest.fixture(autouse=True)
def _scratch_cwd(tmp_path, monkeypatch):
    """run() writes `<fout_name>-<key>_iter-NNNN.{npz,dat}` and `_tmp_wb/` relative to the working directory, as the
    reference does (run_grid.py:244-246,368): every test runs in its own scratch directory."""
    monkeypatch.chdir(tmp_path)


def soc_system_from_fixture(wb, g):
    """the synthetic SOC system of tests/golden/make_golden_soc.py as this package's SystemSOC"""
    def scalar(tag):
        s = wb.System_R(g["real_lattice"], g["iRvec_" + tag], g["centres_" + tag])
        s.set_R_mat("Ham", g["Ham_" + tag])
        s.set_R_mat("AA", g["AA_" + tag])
        return s
    return wb.SystemSOC(scalar("up"), scalar("dw"), iRvec=g["iRvec_soc"], Ham_SOC=g["Ham_SOC"], SS=g["SS"])


SOC_CALCS = dict(dos=("DOS", {}), cumdos=("CumDOS", {}), ahc=("AHC", {}),
                 ahc_int=("AHC", dict(kwargs_formula=dict(external_terms=False))), spin=("Spin", {}),
                 ohmic=("Ohmic_FermiSea", {}), gme_spin=("GME_spin_FermiSurf", {}))
