#!/usr/bin/env python
"""Answer the recorded needle / flash requests of the unmodified reference ON THE GPU (run under gpurun).

tests/golden/dropin_requests/ holds every `needle` / `flash` invocation the reference's run_crispresso made for the four
runs of tests/dropin/runs.py (recorded where /root/reference exists).  This script pipes each of them through the
product executables crispresso_b200/bin/needle and crispresso_b200/bin/flash -- libcrgpu behind the reference's own
command lines -- and stores what they wrote under gpurun_out/dropin_capture/<key>...  Copied to
tests/golden/dropin_capture/, these files are what tests/test_dropin_reference.py replays to the unmodified reference.
"""
import glob
import gzip
import json
import os
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests")]
from dropin import harness  # noqa: E402

REQ = os.path.join(ROOT, "tests", "golden", "dropin_requests")
DATA = os.path.join(ROOT, "tests", "golden", "ref_test_data")
BIN = os.path.join(ROOT, "crispresso_b200", "bin")


def answer_needle(req_path, work):
    key, argv, stdin = harness.load_needle_request(req_path, work)
    env = dict(os.environ, CRGPU_NEEDLE_RUNDATE=harness.RUNDATE)
    p = subprocess.run([sys.executable, os.path.join(BIN, "needle")] + argv, input=stdin, stdout=subprocess.PIPE,
                       stderr=subprocess.PIPE, env=env)
    if p.returncode:
        raise RuntimeError("needle failed on %s: %s" % (key, p.stderr.decode()[-400:]))
    return key, p.stdout


def answer_flash(req_path, work):
    with open(req_path) as f:
        d = json.load(f)
    key = os.path.basename(req_path)[:-len(".req.json")]
    outdir = os.path.join(work, key)
    argv = [{"@R1@": os.path.join(DATA, d["r1"]), "@R2@": os.path.join(DATA, d["r2"]), "@OUTDIR@": outdir}.get(t, t) for t in d["argv"]]
    p = subprocess.run([sys.executable, os.path.join(BIN, "flash")] + argv, stdout=subprocess.PIPE, stderr=subprocess.PIPE)
    if p.returncode:
        raise RuntimeError("flash failed on %s: %s" % (key, p.stderr.decode()[-400:]))
    with gzip.open(os.path.join(outdir, "out.extendedFrags.fastq.gz"), "rb") as f:
        return key, f.read()


def main():
    out = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "gpurun_out", "dropin_capture")
    os.makedirs(out, exist_ok=True)
    work = tempfile.mkdtemp(prefix="dropin_cap_")
    for path in sorted(glob.glob(os.path.join(REQ, "needle_*.req.json.gz"))):
        key, text = answer_needle(path, work)
        with gzip.GzipFile(os.path.join(out, key + ".txt.gz"), "wb", mtime=0) as f:
            f.write(text)
        print(key, len(text), "bytes of srspair text", flush=True)
    for path in sorted(glob.glob(os.path.join(REQ, "flash_*.req.json"))):
        key, data = answer_flash(path, work)
        with gzip.GzipFile(os.path.join(out, key + ".extendedFrags.fastq.gz"), "wb", mtime=0) as f:
            f.write(data)
        print(key, len(data), "bytes of merged FASTQ", flush=True)


if __name__ == "__main__":
    main()
