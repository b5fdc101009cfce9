import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def golden_dir():
    return os.path.join(ROOT, "tests", "golden")


# ---- parity report: every achieved maximum error of the -m gpu parity tests, written at session end ---------------------
_PARITY = {}


@pytest.fixture(scope="session")
def parity_record():
    """record(test, case, field, err, tol): keeps the worst achieved error per (test, case, field)."""

    def record(test, case, field, err, tol=None):
        d = _PARITY.setdefault(test, {}).setdefault(case, {})
        cur = d.get(field)
        if cur is None or err > cur["achieved"]:
            d[field] = {"achieved": float(err), "tolerance": None if tol is None else float(tol)}

    return record


def pytest_sessionfinish(session, exitstatus):
    if not _PARITY:
        return
    import json

    out = os.environ.get("SPHSM_PARITY_REPORT", os.path.join(ROOT, "gpurun_out", "parity_report.json"))
    try:
        os.makedirs(os.path.dirname(out), exist_ok=True)
        with open(out, "w") as fh:
            json.dump({"exitstatus": int(exitstatus), "metric": "max |x - ref| / max(|ref_i|, field scale) (SURVEY.md 8c); "
                       "trajectory entries are absolute world units / millivolts", "tests": _PARITY}, fh, indent=1, sort_keys=True)
    except OSError:
        pass
