import base64
import gzip
import json
import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
if os.path.join(ROOT, "tests") not in sys.path:
    sys.path.insert(0, os.path.join(ROOT, "tests"))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def load_golden():
    """tests/golden/cases.json.gz -- produced by tests/golden/make_golden.py from the reference."""
    path = os.path.join(ROOT, "tests", "golden", "cases.json.gz")
    with gzip.open(path, "rb") as f:
        cases = json.loads(f.read().decode("utf-8"))
    for c in cases:
        c["stdout"] = base64.b64decode(c["stdout_b64"])
    return cases


GOLDEN = load_golden()

# Golden cases on which the reference runs to completion but relies on Python
# behaviour the CUDA path refuses to guess at (DESIGN.md "documented
# deviations"): the product must raise UnsupportedInput, never print a
# different GFA.
UNSUPPORTED_BY_DESIGN = {
    "cs_tilde": "'~' op: the reference reuses a stale length from an earlier op/line",
    "underscore_int": "int('1_5') == 15 in Python; the device integer parser rejects '_'",
    "bare_cr_gaf": "universal-newline translation of a lone CR",
}


@pytest.fixture(scope="session")
def golden_cases():
    return GOLDEN
