"""needle-compatible aligned-pair text (EMBOSS `srspair`, -awidth3=5000) from GPU alignment records.

The reference keeps `needle_output_<id>.txt.gz` when --keep_intermediate / --dump is given
(CRISPResso/CRISPRessoCORE.py:1691, 3694-3697) and re-reads it with parse_needle_output
(CORE:1707-1786).  This writer reproduces the record layout that parser walks:
line offsets relative to "# Aligned_sequences", the Identity line, the 21-column prefix of the
three alignment rows, and the end position it takes as `length`.
"""
import gzip
import time


def fasta_name(fastq_header):
    """Name needle sees: the awk stage keeps the leading '@', sed turns ':' into '_' (CORE:1796-1797);
    EMBOSS takes the first whitespace-delimited token."""
    return ("@" + fastq_header.split()[0]).replace(":", "_")


def _pct(num, den):
    # "%4.1f" of a float32 quotient, as ajalign.c prints it (App. B.3)
    import numpy as np
    return "%4.1f" % float(np.float32(100.0) * np.float32(num) / np.float32(den))


def format_record(aname, bname, ref_row, markup, read_row, ident, score, gapopen, gapextend):
    n = len(ref_row)
    gaps = ref_row.count("-") + read_row.count("-")
    a_end = n - ref_row.count("-")
    b_end = n - read_row.count("-")
    head = (
        "#=======================================\n#\n# Aligned_sequences: 2\n"
        "# 1: %s\n# 2: %s\n# Matrix: EDNAFULL\n# Gap_penalty: %.1f\n# Extend_penalty: %.1f\n#\n"
        "# Length: %d\n"
        "# Identity:   %5d/%d (%s%%)\n# Similarity: %5d/%d (%s%%)\n# Gaps:       %5d/%d (%s%%)\n"
        "# Score: %.1f\n# \n#\n#=======================================\n\n"
    ) % (aname, bname, gapopen, gapextend, n, ident, n, _pct(ident, n), ident, n, _pct(ident, n), gaps, n, _pct(gaps, n), score)
    rows = (
        "%-13.13s %6d %s %6d\n" % (aname, 1 if a_end else 0, ref_row, a_end)
        + " " * 21 + markup + "\n"
        + "%-13.13s %6d %s %6d\n" % (bname, 1 if b_end else 0, read_row, b_end)
    )
    return head + rows + "\n\n"


FILE_TRAILER = "#---------------------------------------\n#---------------------------------------\n"


def _num(x):
    return ("%g" % x) if isinstance(x, float) else str(x)


def file_header(asequence="amplicon.fa", bsequence="/dev/stdin", outfile="/dev/stdout", gapopen=10.0, gapextend=0.5,
                awidth3=5000, rundate=None):
    """The block needle writes once per run (ignored by parse_needle_output).  `rundate` is the only part of the
    output that is not a function of the inputs; CRGPU_NEEDLE_RUNDATE pins it (reproducible captures)."""
    import os
    rundate = rundate or os.environ.get("CRGPU_NEEDLE_RUNDATE") or time.strftime("%a %e %b %Y %H:%M:%S")
    aw = "" if awidth3 is None else "#    -awidth3=%s\n" % awidth3
    return ("########################################\n# Program: needle\n# Rundate: %s\n"
            "# Commandline: needle\n#    -asequence=%s\n#    -bsequence=%s\n#    -outfile=%s\n"
            "#    -gapopen=%s\n#    -gapextend=%s\n%s# Align_format: srspair\n"
            "# Report_file: %s\n########################################\n\n"
            % (rundate, asequence, bsequence, outfile, _num(gapopen), _num(gapextend), aw, outfile))


def write_needle_output(path, aname, bnames, recs, ref_rows, markup_rows, read_rows, gapopen=10.0, gapextend=0.5,
                        asequence="amplicon.fa"):
    """Write one srspair record per alignment (gzip if path ends with .gz)."""
    op = gzip.open if path.endswith(".gz") else open
    with op(path, "wt") as f:
        f.write(file_header(asequence, gapopen=gapopen, gapextend=gapextend))
        for i, b in enumerate(bnames):
            f.write(format_record(aname, b, ref_rows[i], markup_rows[i], read_rows[i], int(recs["ident"][i]),
                                  float(recs["score"][i]), gapopen, gapextend))
        f.write(FILE_TRAILER)
