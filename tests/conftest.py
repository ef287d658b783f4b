"""pytest configuration: the `gpu` marker and shared fixtures.

`-m "not gpu"` (CPU, this container and CI): oracle vs the reference's goldens, host logic,
C-ABI surface.  `-m gpu` (a B200): parity of the CUDA path against the oracle, through the
C-ABI.  Nothing here reads /root/reference at run time.
"""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
sys.path.insert(0, os.path.join(ROOT, "one-class-ffm_b200"))
GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def load_golden(name):
    z = np.load(os.path.join(GOLDEN, name + ".npz"))
    d = {k: z[k] for k in z.files}
    d["flags"] = bytes(d["meta.flags"]).decode().split()
    return d


def params_of(d):
    lam, omega, r, k, nr_pass, self_side, freq = d["params"]
    return dict(k=int(k), lam=float(lam), omega=float(omega), r=float(r),
                self_side=bool(self_side), freq=bool(freq)), int(nr_pass)


@pytest.fixture(scope="session", params=["tiny", "tiny_ns", "tiny_freq", "small"])
def golden(request):
    return request.param, load_golden(request.param)
