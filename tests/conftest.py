import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with `-m gpu`)")


@pytest.fixture(scope="session")
def lib():
    from climaatmos_jl_b200 import capi

    capi.build()
    return capi.load()


def record_parity(name, row):
    """Append one measured-parity row to gpurun_out/parity_report.jsonl (evidence copied under profiles/ per round); never fails a test."""
    import json

    try:
        out = os.path.join(ROOT, "gpurun_out")
        os.makedirs(out, exist_ok=True)
        with open(os.path.join(out, "parity_report.jsonl"), "a") as f:
            f.write(json.dumps(dict(test=name, **row)) + "\n")
    except Exception:
        pass
