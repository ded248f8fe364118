"""The UNMODIFIED reference run_crispresso (CRISPResso/CRISPRessoCORE.py:1216) over the drop-in executables.

Build container only (needs /root/reference; skipped elsewhere).  `replay` runs feed the reference the bytes the
GPU executables produced on a B200 (tests/golden/dropin_capture/, made by scripts/make_dropin_capture.py; the GPU
test tests/test_gpu_dropin.py checks that the capture is current): the reference's own parse_needle_output
(CORE:1707-1786), ID round trip (CORE:1725), merge / RC-rescue plumbing (CORE:1830-2010), process_df_chunk
(CORE:428-753) and reductions consume GPU output and must return the golden 14-tuple of its own tests
(tests/crispresso_tests.py:181-195), with -p 1 and -p 2.
"""
import glob
import os

import pytest

import ref_shim
from dropin import fused, harness, runs

pytestmark = pytest.mark.skipif(not ref_shim.available(), reason="the reference tree only exists in the build container")
CAPTURE = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "dropin_capture")


def _replay_bin(tmp_path):
    if not glob.glob(os.path.join(CAPTURE, "needle_*.txt.gz")):
        pytest.fail("tests/golden/dropin_capture/ is empty: run scripts/make_dropin_capture.py under gpurun and commit the result")
    os.environ["CRGPU_DROPIN_REPLAY"] = CAPTURE
    return harness.make_bin_dir(str(tmp_path / "bin"), "replay")


def test_reference_kat1_on_gpu_output_two_processes(tmp_path):
    """tests/crispresso_tests.py:127-195 with p = 2, keep_intermediate = True."""
    out = runs.run("kat1", _replay_bin(tmp_path), str(tmp_path / "out"), n_processes=2, keep_intermediate=True)
    fused.assert_matches_reference("kat1", runs.summarize(out))
    assert int(out[1]) == 8906                                   # n_reads_input
    kept = glob.glob(str(tmp_path / "out" / "*" / "needle_output_*.txt.gz"))
    assert kept, "needle_output_<id>.txt.gz is kept with --keep_intermediate (CORE:3694-3697)"


def test_reference_kat2_untrimmed_on_gpu_output(tmp_path):
    """tests/crispresso_tests.py:198-272 without Trimmomatic: --min_identity_score 30, -w 23, reverse-complement rescue hits."""
    out = runs.run("kat2_untrimmed", _replay_bin(tmp_path), str(tmp_path / "out"), n_processes=1)
    fused.assert_matches_reference("kat2_untrimmed", runs.summarize(out))


@pytest.mark.parametrize("name", ["cfg1_single_end", "hdr_coding"])
def test_reference_small_runs_on_gpu_output(tmp_path, name):
    out = runs.run(name, _replay_bin(tmp_path), str(tmp_path / "out"), n_processes=1)
    got = runs.summarize(out)
    exp = fused.expected()[name]
    assert {k: exp[k] for k in got} == got


def test_reference_cfg1_over_the_oracle_front_end(tmp_path):
    """Same executables' front ends (argument parsing, FASTA reading, srspair writer, FLASH file layout) with the CPU
    oracle behind them: what the recorded requests and expected.json were made with."""
    bin_dir = harness.make_bin_dir(str(tmp_path / "bin"), "oracle")
    out = runs.run("cfg1_single_end", bin_dir, str(tmp_path / "out"), n_processes=2)
    got = runs.summarize(out)
    exp = fused.expected()["cfg1_single_end"]
    assert {k: exp[k] for k in got} == got


@pytest.mark.parametrize("name", ["kat1", "kat2_untrimmed", "cfg1_single_end", "hdr_coding"])
def test_oracle_fused_path_returns_what_the_reference_run_returned(name):
    """Pins oracle/quantify.py's restatement of CORE:1830-2072 + process_df_chunk + postreduce against what the real
    reference code returned for the same reads, including the HDR / MIXED / frameshift branches its own tests never reach."""
    fused.assert_matches_reference(name, fused.summarize_oracle(name))
