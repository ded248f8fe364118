#!/bin/bash
# compute-sanitizer evidence for a round (run under gpurun): memcheck, racecheck, initcheck and synccheck over
# scripts/gpu_sanitize_cases.py.   usage: scripts/sanitize_round.sh <tag>  -> gpurun_out/<tag>_sanitizer_<tool>.txt
tag=${1:-rXX}
out=gpurun_out
mkdir -p $out
python scripts/gpu_sanitize_cases.py > $out/${tag}_sanitizer_plain.txt 2>&1 || { tail -5 $out/${tag}_sanitizer_plain.txt; exit 1; }
for tool in memcheck racecheck synccheck initcheck; do
    extra=""
    [ $tool = memcheck ] && extra="--leak-check full"
    [ $tool = initcheck ] && extra="--track-unused-memory no"
    timeout 1500 compute-sanitizer --tool $tool $extra --error-exitcode 9 --print-limit 20 python scripts/gpu_sanitize_cases.py > $out/${tag}_sanitizer_$tool.txt 2>&1
    echo "$tool: exit $? -- $(grep -E 'ERROR SUMMARY|RACECHECK SUMMARY|LEAK SUMMARY' $out/${tag}_sanitizer_$tool.txt | tr '\n' ' ')"
done
