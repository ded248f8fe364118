"""Short runs of the randomised differential tests (scripts/gpu_fuzz_*.py): random amplicon lengths
2..1024 (every kernel tile), read lengths 2..2048, unrelated / shifted / edited reads, N, ten dyadic
gap-penalty pairs; and random CRISPResso option sets for the fused path.  Longer runs of the same
scripts (5384 + 5164 cases, 0 mismatches) are recorded in profiles/r01_notes.md.  Needs a B200."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("script,seed", [("gpu_fuzz_align.py", 11), ("gpu_fuzz_hotpath.py", 12)])
def test_fuzz(script, seed):
    r = subprocess.run([sys.executable, os.path.join(ROOT, "scripts", script), str(seed), "10"], cwd=ROOT,
                       capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    assert "0 mismatching cases" in r.stdout
