"""Loader for the committed golden fixtures (tests/golden/)."""
import gzip
import json
import os

HERE = os.path.dirname(os.path.abspath(__file__))


def load_process_df_chunk_cases():
    with gzip.open(os.path.join(HERE, "golden", "process_df_chunk_cases.json.gz"), "rt") as f:
        data = json.load(f)
    for c in data["cases"]:
        for r in c["rows"]:
            if "score_repaired" in r and r["score_repaired"] is None:
                r["score_repaired"] = float("nan")
    return data["cases"]
