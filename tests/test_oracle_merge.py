"""The FLASH restatement (oracle/flash_merge.py) against the committed pair fixture, and the host-side
pieces of crispresso_b200/flash.py that need no GPU."""
import gzip
import json
import os

from crispresso_b200 import flash
from oracle import flash_merge

HERE = os.path.dirname(os.path.abspath(__file__))


def test_oracle_reproduces_the_fixture():
    with gzip.open(os.path.join(HERE, "golden", "flash_pairs_subset.json.gz"), "rt") as f:
        pairs = json.load(f)["pairs"]
    kinds = set()
    for p in pairs[::5]:
        m = flash_merge.merge_pair(p["s1"], p["q1"], p["s2"], p["q2"])
        assert (list(m) if m else None) == p["merged"]
        kinds.add(m[2] if m else None)
    assert kinds == {"innie", "outie", None}


def test_combined_tag():
    assert flash.combined_tag("M1:1:2 1:N:0:1", "M1:1:2 2:N:0:1") == "M1:1:2 1:N:0:1"
    assert flash.combined_tag("read7/1", "read7/2") == "read7"
    assert flash.combined_tag("read7", "read7") == "read7"
