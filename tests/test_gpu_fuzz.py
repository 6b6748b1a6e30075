"""Device-side fuzzing of the page readers: mutated page headers / level blocks / index streams must be rejected with
HS_EFORMAT (or decode harmlessly), never cause an out-of-bounds access (tests/fuzz_pages.py, own process)."""
import json
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_mutated_pages_never_fault_the_device():
    r = subprocess.run([sys.executable, os.path.join(ROOT, "tests", "fuzz_pages.py"), "--iterations", "400", "--seed", "7"],
                       capture_output=True, text=True, timeout=900, cwd=ROOT)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-3000:]
    rep = json.loads(r.stdout.strip().splitlines()[-1])
    assert rep["counts"]["cuda_error"] == 0 and rep["healthy"]
    assert rep["counts"]["rejected"] > 20, rep  # the mutations do reach the validation code
    assert rep["counts"]["ok"] + rep["counts"]["rejected"] == rep["iterations"]
