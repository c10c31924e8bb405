import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")
    config.addinivalue_line(
        "markers", "needs_reference: needs /root/reference (authoring container only; auto-skipped elsewhere)"
    )


def pytest_collection_modifyitems(config, items):
    have_ref = os.path.isdir("/root/reference/cnvlib")
    skip_ref = pytest.mark.skip(reason="/root/reference not mounted")
    for item in items:
        if "needs_reference" in item.keywords and not have_ref:
            item.add_marker(skip_ref)


def load_golden(name):
    return np.load(os.path.join(GOLDEN, name + ".npz"))


@pytest.fixture(scope="session")
def golden():
    return load_golden


def golden_table(g, prefix):
    """Rebuild a CopyNumArray from the 'prefix__column' arrays written by oracle/make_golden.py."""
    import pandas as pd

    from cnvkit_b200.cnary import CopyNumArray

    cols = [str(c) for c in g[f"{prefix}__columns"]]
    data = {}
    for c in cols:
        v = g[f"{prefix}__{c}"]
        data[c] = v.astype(object) if v.dtype.kind == "U" else v
    return CopyNumArray(pd.DataFrame(data), {"sample_id": prefix})
