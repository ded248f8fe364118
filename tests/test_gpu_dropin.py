"""The drop-in executables on the B200: crispresso_b200/bin/needle and crispresso_b200/bin/flash answer the very
command lines the UNMODIFIED reference issued (recorded under tests/golden/dropin_requests/ by running its
run_crispresso, CORE:1216, in the build container), and the fused path reproduces the numbers that reference run
returned.  Needs a B200."""
import glob
import gzip
import io
import os
import sys
import tempfile

import pytest

from dropin import fused, harness

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "scripts"))
CAPTURE = os.path.join(ROOT, "tests", "golden", "dropin_capture")


@pytest.fixture(scope="module")
def answers():
    """Every recorded request answered by the product executables (subprocesses, as the reference's shell runs them)."""
    import make_dropin_capture as cap
    work = tempfile.mkdtemp(prefix="dropin_gpu_")
    out = {}
    for path in sorted(glob.glob(os.path.join(fused.REQ, "needle_*.req.json.gz"))):
        key, text = cap.answer_needle(path, work)
        out[key] = text
    for path in sorted(glob.glob(os.path.join(fused.REQ, "flash_*.req.json"))):
        key, data = cap.answer_flash(path, work)
        out[key] = data
    return out, work


def test_needle_executable_writes_what_the_oracle_writes(answers):
    """Byte for byte: the srspair text of the GPU `needle` == the same front end over the CPU oracle, for every needle
    invocation of the four reference runs (forward, HDR, reverse-complement rescue)."""
    from crispresso_b200 import needle_cli
    got, work = answers
    n = 0
    for path in sorted(glob.glob(os.path.join(fused.REQ, "needle_*.req.json.gz"))):
        key, argv, stdin = harness.load_needle_request(path, work)
        os.environ["CRGPU_NEEDLE_RUNDATE"] = harness.RUNDATE
        buf, real = io.StringIO(), sys.stdout
        sys.stdout = buf
        try:
            assert needle_cli.main(argv, align=harness.oracle_align, stdin=io.BytesIO(stdin)) == 0
        finally:
            sys.stdout = real
        assert got[key] == buf.getvalue().encode(), key
        n += 1
    assert n >= 9


def test_flash_executable_writes_what_the_oracle_writes(answers):
    import json

    from crispresso_b200 import flash_cli
    got, work = answers
    for path in sorted(glob.glob(os.path.join(fused.REQ, "flash_*.req.json"))):
        with open(path) as f:
            d = json.load(f)
        key = os.path.basename(path)[:-len(".req.json")]
        outdir = os.path.join(work, key + "_oracle")
        os.makedirs(outdir)
        data = os.path.join(ROOT, "tests", "golden", "ref_test_data")
        argv = [{"@R1@": os.path.join(data, d["r1"]), "@R2@": os.path.join(data, d["r2"]), "@OUTDIR@": outdir}.get(t, t) for t in d["argv"]]
        assert flash_cli.main(argv, merge=harness.oracle_merge) == 0
        with gzip.open(os.path.join(outdir, "out.extendedFrags.fastq.gz"), "rb") as f:
            assert got[key] == f.read(), key


def test_committed_capture_is_what_the_gpu_writes_now(answers):
    """tests/golden/dropin_capture/ (what the CPU-side replay test feeds to the unmodified reference) is current."""
    got, _work = answers
    files = glob.glob(os.path.join(CAPTURE, "*.gz"))
    assert files, "no capture committed: run scripts/make_dropin_capture.py under gpurun"
    for path in files:
        key = os.path.basename(path).split(".")[0]
        with gzip.open(path, "rb") as f:
            assert got[key] == f.read(), key


@pytest.mark.parametrize("name", ["kat1", "kat2_untrimmed", "cfg1_single_end", "hdr_coding"])
def test_fused_path_returns_what_the_reference_run_returned(ctx, name):
    """crgpu_align_quantify + postreduce on the reads the reference aligned == the 14-tuple its run_crispresso returned
    (CORE:3977-3992): KAT #1 (all golden values), KAT #2's trimming-independent values, cfg1 (single end,
    --min_identity_score 50), HDR + coding sequence."""
    got, _res = fused.summarize_gpu(ctx, name)
    fused.assert_matches_reference(name, got)


def test_a_real_needle_on_the_box_agrees(answers):
    """BASELINE.md section 4 step 1 / SURVEY hard part 6: if an EMBOSS needle is installed on the GPU box, its output for the
    recorded requests must equal ours line for line (header block aside)."""
    import shutil
    import subprocess
    real = shutil.which("needle")
    if real is None or os.path.dirname(real) == os.path.join(ROOT, "crispresso_b200", "bin"):
        pytest.skip("no EMBOSS needle on this box (which needle: %s)" % real)
    got, work = answers
    for path in sorted(glob.glob(os.path.join(fused.REQ, "needle_*.req.json.gz"))):
        key, argv, stdin = harness.load_needle_request(path, work)
        p = subprocess.run([real] + argv, input=stdin, stdout=subprocess.PIPE, stderr=subprocess.PIPE)
        assert p.returncode == 0

        def body(b):
            return [ln for ln in b.decode().split("\n") if not ln.startswith(("# Rundate", "#    -", "# Commandline", "# Report_file"))]
        assert body(p.stdout) == body(got[key]), key
