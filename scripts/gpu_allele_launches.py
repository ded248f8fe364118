"""One staged crgpu_align_quantify call with the allele table on cfg2's reads: run under
ncu --metrics gpu__time_duration.sum to list the allele kernels (scripts/profile_round.sh does)."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from crispresso_b200 import Context, _lib, hotpath  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 20
ctx = Context(0)
jobs, _ = bench.workload("cfg2", n, 0)
j = jobs[0]
res = hotpath.run_hot_path_staged(ctx, j.amp, (j.buf, j.off), chunk_reads=n, alleles=1 << 16, **j.kw)
print("alleles:", len(res.allele_count) if hasattr(res, "allele_count") else res)
