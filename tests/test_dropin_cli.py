"""Front ends of the drop-in executables (crispresso_b200/needle_cli.py, flash_cli.py) -- CPU only: argument
handling as the reference's command lines need it (CORE:1655-1664, 1791-1936); the alignments come from the oracle."""
import io
import sys

import pytest

from crispresso_b200 import flash_cli, needle_cli, synth
from dropin import harness


def test_needle_accepts_the_reference_command_line_and_refuses_the_malformed_one(tmp_path):
    ok = "-asequence=db.fa -bsequence=/dev/stdin -outfile=/dev/stdout -gapopen=10 -gapextend=0.5  -awidth3=5000".split()
    opts, go, ge = needle_cli.parse_command_line(ok)
    assert (go, ge) == (10.0, 0.5) and opts["asequence"] == "db.fa"
    # CORE:1924-1936: the repair-RC command carries the literal text instead of the options -> needle dies, no record
    bad = "-asequence=db.fa -bsequence=/dev/stdin -outfile=/dev/stdout args.needle_options_string".split()
    with pytest.raises(needle_cli.UsageError):
        needle_cli.parse_command_line(bad)
    assert needle_cli.main(bad, align=harness.oracle_align, stdin=io.BytesIO(b">x\nACGT\n")) == 1
    with pytest.raises(needle_cli.UsageError):
        needle_cli.parse_command_line(ok + ["-datafile=EBLOSUM62"])       # options the aligner cannot honour are refused


def test_needle_front_end_round_trip(tmp_path, capsys):
    """FASTA in (names = first token, gaps dropped as EMBOSS does for the RC-rescue input, CORE:1846) -> srspair out."""
    amp, _g, cut, _h = synth.make_case(3, 120, hdr=False)
    fa = tmp_path / "db.fa"
    fa.write_text(">amp1\n%s\n" % amp)
    reads = [amp, amp[:50] + amp[58:], amp[:60] + "TTT" + amp[60:]]
    gapped = reads[1][:30] + "--" + reads[1][30:]
    stdin = "".join(">@r_%d extra words\n%s\n" % (i, s) for i, s in enumerate(reads + [gapped])).encode()
    rc = needle_cli.main(["-asequence=%s" % fa, "-bsequence=/dev/stdin", "-outfile=/dev/stdout", "-gapopen=10", "-gapextend=0.5",
                          "-awidth3=5000"], align=harness.oracle_align, stdin=io.BytesIO(stdin))
    assert rc == 0
    text = capsys.readouterr().out
    assert text.count("# Aligned_sequences: 2") == 4 and "# 2: @r_3\n" in text and "# 1: amp1\n" in text
    recs = text.split("#=======================================\n#\n# Aligned_sequences")[1:]
    assert "(100.0%)" in recs[0]
    assert recs[1].split("\n\n")[1] == recs[3].split("\n\n")[1].replace("@r_3", "@r_1")      # the gapped copy aligns like the read


def test_flash_accepts_the_reference_command_line():
    argv = "a_R1.fastq.gz a_R2.fastq.gz --allow-outies --max-overlap 100 --min-overlap 4 -f 280 -r 151 -s 28  -z -d /tmp/x".split()
    files, o = flash_cli.parse_command_line(argv)
    assert files == ["a_R1.fastq.gz", "a_R2.fastq.gz"]
    assert (o["min_overlap"], o["max_overlap"], o["allow_outies"], o["gz"], o["outdir"]) == (4, 100, True, True, "/tmp/x")
    with pytest.raises(flash_cli.UsageError):
        flash_cli.parse_command_line(argv + ["--interleaved-input"])
