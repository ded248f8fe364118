"""oracle/quantify.py::process_rows against vectors produced by the UNMODIFIED reference
process_df_chunk (tests/golden/make_golden.py).  CPU only."""
import numpy as np
import pytest

from golden_io import load_process_df_chunk_cases
from oracle import quantify

CASES = load_process_df_chunk_cases()


@pytest.mark.parametrize("case", CASES, ids=[c["name"] for c in CASES])
def test_process_rows_matches_reference(case):
    opts = quantify.Opts(**case["opts"])
    per_row, V, hin, hfs, cnt = quantify.process_rows(case["rows"], opts, case["include"], case["L"], case["exon"],
                                                      case["splice"])
    exp = case["expected"]
    assert per_row == exp["per_row"]
    for k in quantify.VECTOR_NAMES:
        assert V[k].tolist() == exp["vectors"][k], k
    assert {str(k): v for k, v in hin.items()} == exp["hist_inframe"]
    assert {str(k): v for k, v in hfs.items()} == exp["hist_frameshift"]
    assert cnt == exp["counters"]


def test_ref_positions_and_revcomp():
    assert quantify.ref_positions("--AC-GT--") == [-1, -1, 0, 1, -2, 2, 3, -4, -4]
    assert quantify.reverse_complement("ACTGGT") == "ACCAGT"      # reference tests/crispresso_tests.py:99-101
    assert quantify.mask_n("ANGT", "|.||") == ("||||", True)
