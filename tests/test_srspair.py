"""The needle-compatible text writer against a restatement of the reference's parser
(CRISPResso/CRISPRessoCORE.py:1707-1786, a closure inside run_crispresso and therefore not importable).
CPU only: alignments come from the oracle."""
import gzip

import numpy as np

from crispresso_b200 import srspair, synth
from oracle import needle


def parse_needle_output(path, just_score=False):
    """Line-for-line semantics of CORE:1715-1765."""
    out = []
    with gzip.open(path, "r") as f:
        line = f.readline().decode()
        while line:
            while line and "# Aligned_sequences" not in line:
                line = f.readline().decode()
            if line:
                f.readline()
                line = f.readline().decode()
                id_seq = line.split()[-1].replace("_", ":")
                for _ in range(5):
                    f.readline()
                line = f.readline().decode()
                ident = float(line.strip().split(" ")[-1].replace("%", "").replace(")", "").replace("(", ""))
                if just_score:
                    out.append([id_seq, ident])
                else:
                    for _ in range(7):
                        f.readline()
                    line = f.readline().decode()
                    ref = line.split()[2]
                    mark = f.readline().decode()[21:].rstrip("\n")
                    line = f.readline().decode()
                    out.append([id_seq, ident, line.split()[3], ref, mark, line.split()[2]])
    return out


def test_written_text_parses_back_to_the_same_columns(tmp_path):
    amp, _g, cut, hdr = synth.make_case(31, 180)
    buf, off = synth.make_reads(amp, hdr, cut, 120, seed=31)
    reads = [bytes(buf[off[i]:off[i + 1]]).decode() for i in range(120)] + [amp[:40], "ACGT" * 20]
    res, ref, mark, qry = needle.align_batch(amp, reads)
    names = [srspair.fasta_name("M06879:15:000000000-DFF22:1:1101:%d:%d 1:N:0:1" % (i, 7 * i)) for i in range(len(reads))]
    recs = np.zeros(len(reads), dtype=[("ident", "<i4"), ("score", "<f4")])
    recs["ident"], recs["score"] = res["ident"], res["score"]
    p = str(tmp_path / "needle_output_x.txt.gz")
    srspair.write_needle_output(p, "x", names, recs, ref, mark, qry)
    rows = parse_needle_output(p)
    assert len(rows) == len(reads)
    for i, (id_seq, ident, length, r, m, q) in enumerate(rows):
        assert id_seq == "@M06879:15:000000000-DFF22:1:1101:%d:%d" % (i, 7 * i)
        assert ident == res["tenths"][i] / 10.0
        assert length == str(len(reads[i]))
        assert (r, m, q) == (ref[i], mark[i], qry[i])
    assert [x[1] for x in parse_needle_output(p, just_score=True)] == [x[1] for x in rows]


def test_record_layout_matches_the_documented_example():
    txt = srspair.format_record("amp", "read_1", "ACGT-ACGT", "|||| .|||", "ACGTTTCGT", 7, 20.5, 10.0, 0.5)
    lines = txt.split("\n")
    assert lines[2] == "# Aligned_sequences: 2" and lines[4] == "# 2: read_1"
    assert lines[10] == "# Identity:       7/9 (77.8%)"
    assert lines[12] == "# Gaps:           1/9 (11.1%)"
    assert lines[13] == "# Score: 20.5"
    assert lines[18] == "amp                1 ACGT-ACGT      8"
    assert lines[19] == " " * 21 + "|||| .|||"
    assert lines[20] == "read_1             1 ACGTTTCGT      9"
