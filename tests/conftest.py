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
