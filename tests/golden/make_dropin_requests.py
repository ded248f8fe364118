"""Record what the UNMODIFIED reference asks of `needle` and `flash`, and what it returns.

Runs CRISPRessoCORE.run_crispresso (CORE:1216, imported from /root/reference through tests/ref_shim.py) over the
`oracle` flavour of tests/dropin/harness.py for four inputs and writes
  tests/golden/dropin_requests/<key>.req.json[.gz]   every needle / flash invocation (command line + inputs),
  tests/golden/dropin_requests/expected.json         the 14-tuple each run returned (scalars + histogram heads),
  tests/golden/ref_test_data/sub_*.fastq.gz          the first 3000 records of the test_L001 pair (smaller runs).
The requests are answered on a B200 by scripts/make_dropin_capture.py (-> tests/golden/dropin_capture/), and
tests/test_dropin_reference.py feeds those GPU answers back to the unmodified reference.
Build container only (needs /root/reference).
"""
import gzip
import json
import os
import shutil
import sys
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path[:0] = [os.path.join(HERE, "..", ".."), os.path.join(HERE, "..")]
from dropin import harness, runs  # noqa: E402


def subsample(src, dst, n):
    with gzip.open(src, "rt") as f, gzip.GzipFile(dst, "wb", mtime=0) as g:
        for _ in range(4 * n):
            g.write(f.readline().encode())


def spike_hdr(src, dst):
    """Every 20th read 1 that carries the amplicon's bases at runs.HDR_POS gets the HDR product's bases there; every
    20th after those also gets eight more substitutions (an imperfect HDR: the MIXED class).  (The reference's HDR /
    MIXED plots divide by zero when fewer than six reads fall in a class, CORE:3091, 3203.)"""
    import kat_common as K
    lo, hi = runs.HDR_POS
    with gzip.open(src, "rt") as f:
        lines = f.read().split("\n")
    n = 0
    for i in range(1, len(lines), 4):
        seq = lines[i]
        if (i // 4) % 20 == 0 and len(seq) >= hi + 10 and seq[lo - 20:hi + 10] == K.AMPLICON[lo - 20:hi + 10]:
            lines[i] = seq[:lo] + runs.HDR_AMPLICON[lo:hi] + seq[hi:]
            n += 1
        elif (i // 4) % 20 == 10 and len(seq) >= hi + 10 and seq[lo - 20:hi + 10] == K.AMPLICON[lo - 20:hi + 10]:
            sub = "".join({"A": "C", "C": "G", "G": "T", "T": "A", "N": "N"}[c] for c in seq[30:38])
            lines[i] = seq[:30] + sub + seq[38:lo] + runs.HDR_AMPLICON[lo:hi] + seq[hi:]
            n += 1
    with gzip.GzipFile(dst, "wb", mtime=0) as g:
        g.write("\n".join(lines).encode())
    return n


def main():
    data = os.path.join(HERE, "ref_test_data")
    for name in ("test_L001_R1_001", "test_L001_R2_001"):
        subsample(os.path.join(data, name + ".fastq.gz"), os.path.join(data, "sub_" + name + ".fastq.gz"), 3000)
    print("HDR reads written:", spike_hdr(os.path.join(data, "sub_test_L001_R1_001.fastq.gz"),
                                          os.path.join(data, "hdr_sub_test_L001_R1_001.fastq.gz")))
    req = os.path.join(HERE, "dropin_requests")
    if len(sys.argv) == 1:
        shutil.rmtree(req, ignore_errors=True)
    os.makedirs(req, exist_ok=True)
    os.environ["CRGPU_DROPIN_REQUESTS"] = req
    work = tempfile.mkdtemp(prefix="dropin_req_")
    bin_dir = harness.make_bin_dir(os.path.join(work, "bin"), "oracle")
    expected = {}
    only = sys.argv[1:]
    if only:                                   # re-record some runs: keep the others' requests and expectations
        with open(os.path.join(req, "expected.json")) as f:
            expected = json.load(f)
    log = os.path.join(req, "order.log")
    for name in (only or runs.RUNS):
        if os.path.exists(log):
            os.remove(log)
        out = runs.run(name, bin_dir, os.path.join(work, name))
        expected[name] = runs.summarize(out)
        with open(log) as f:
            expected[name]["requests"] = f.read().split()          # flash (paired end), then the needle calls in order
        os.remove(log)
        print(name, expected[name], flush=True)
    with open(os.path.join(req, "expected.json"), "wt") as f:
        json.dump(expected, f, indent=1, sort_keys=True)
    shutil.rmtree(work, ignore_errors=True)


if __name__ == "__main__":
    main()
