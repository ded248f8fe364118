#!/bin/bash
# Collect the round's evidence on a B200 (run under gpurun): bench line, ncu launch list, ncu --set full of the
# dominant kernels.  usage: scripts/profile_round.sh <tag>   -> gpurun_out/<tag>_*
set -u
tag=${1:-rXX}
out=gpurun_out
mkdir -p $out
python bench.py --steps 5 --warmup 3 > $out/${tag}_bench_1gpu.json 2> $out/${tag}_bench.err || { tail -5 $out/${tag}_bench.err; exit 1; }
# launch list (cold-cache, serialised: compare SHARES with the live CUDA-event figures of the bench line)
ncu --metrics gpu__time_duration.sum --clock-control none -c 1200 --csv --log-file $out/${tag}_launches.csv \
    python bench.py --steps 1 --warmup 1 --no-cpu-baseline > $out/${tag}_ncu_list.log 2>&1
# full sections: two batches of the banded fill + their walks, then the quantifier and the single-pass fill (escapes)
ncu --set full --clock-control none --import-source on -k regex:"k_gotoh_score|k_gotoh_band|k_traceback_walk|k_diag_emit" \
    -s 6 -c 12 -o $out/${tag}_kernels -f python bench.py --steps 1 --warmup 1 --reads 1048576 --no-cpu-baseline > $out/${tag}_ncu_full.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"k_quantify|k_gotoh_fill" \
    -c 3 -o $out/${tag}_kernels2 -f python bench.py --steps 1 --warmup 1 --reads 1048576 --no-cpu-baseline >> $out/${tag}_ncu_full.log 2>&1
tail -2 $out/${tag}_ncu_full.log
# gpurun copies back at most 64 MiB: keep the CSV pages, drop the reports
for r in kernels kernels2; do
    ncu -i $out/${tag}_$r.ncu-rep --page raw --csv > $out/${tag}_${r}_raw.csv 2>/dev/null
    ncu -i $out/${tag}_$r.ncu-rep --page source --csv > $out/${tag}_${r}_source.csv 2>/dev/null
    rm -f $out/${tag}_$r.ncu-rep
done
ls -la $out | grep ${tag}
