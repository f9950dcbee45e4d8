"""CPU: the vote kernel's guard band checked empirically.  tools/band_check.c re-states the fast
cone test of csrc/vote.cu next to the reference predicate and searches boundary-concentrated samples for
disagreements that the band would NOT flag (exit code 1) -- the full 1e8-sample runs are in profiles/."""
import os
import re
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def band_check(tmp_path_factory):
    exe = str(tmp_path_factory.mktemp("band") / "band_check")
    cc = os.environ.get("CC", "gcc")
    flags = ["-O2", "-ffp-contract=off"]
    with open("/proc/cpuinfo") as fh:
        if " fma " in fh.read():
            flags.append("-mfma")
    subprocess.check_call([cc] + flags + [os.path.join(ROOT, "tools", "band_check.c"), "-lm", "-o", exe])
    return exe


@pytest.mark.parametrize("thresh,local", [("0.99", "1"), ("0.999", "1"), ("0.9", "1"), ("0.5", "1"), ("0.99", "0")])
def test_no_unflagged_disagreement(band_check, thresh, local):
    r = subprocess.run([band_check, "3000000", thresh, local], stdout=subprocess.PIPE, text=True, timeout=120)
    assert r.returncode == 0, r.stdout
    mism = int(re.search(r"mismatches=(\d+)", r.stdout).group(1))
    ratio = float(re.search(r"analytic bound .* = ([0-9.]+)", r.stdout).group(1))
    assert mism > 1000            # the sampler really sits on the decision boundary
    assert ratio < 0.8            # worst disagreement stays well inside the analytic bound (band = 1.25 x bound)
