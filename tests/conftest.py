import glob
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN_DIR = os.path.join(ROOT, "tests", "golden")
DATA_DIR = os.path.join(ROOT, "data")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run with -m gpu on the B200 box)")


def golden_names():
    return sorted(os.path.splitext(os.path.basename(p))[0] for p in glob.glob(os.path.join(GOLDEN_DIR, "*.npz")))


def load_golden(name):
    z = np.load(os.path.join(GOLDEN_DIR, name + ".npz"))
    d = {k: z[k] for k in z.files}
    for k in ("n_train", "n_test", "g", "m", "T", "max_iters"):
        d[k] = int(d[k])
    for k in ("approx", "skip_variance"):
        d[k] = bool(d[k])
    d["delta"] = float(d["delta"])
    off = d["offsets"]
    seqs = [d["codes"][off[i]:off[i + 1]].tolist() for i in range(len(off) - 1)]
    d["Xtrain"], d["Xtest"] = seqs[:d["n_train"]], seqs[d["n_train"]:]
    return d


@pytest.fixture(scope="session")
def oracle_mod():
    import oracle
    oracle.c_lib()
    return oracle
