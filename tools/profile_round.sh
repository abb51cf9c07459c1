#!/bin/bash
# Usage (under gpurun): bash tools/profile_round.sh <tag>
# 1) launch list of the bench command (gpu__time_duration per launch, cold-cache & serialised: compare SHARES)
# 2) one --set full capture of the top kernels
TAG=${1:-r01}
mkdir -p gpurun_out
KREGEX='regex:(resize_level|fast_cells|quadtree_kernel|prep_rowstat|prep_map|prep_imgstat|match_prepare|describe_kernel|finalize_kernel|georef_kernel|match_pair_kernel|scc_merge|scan_counts|emit_rows)'
CMD="python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline --no-bruteforce"
ncu --metrics gpu__time_duration.sum --clock-control none -k "$KREGEX" -c 4000 --csv --log-file gpurun_out/launches_${TAG}.csv $CMD > gpurun_out/launches_${TAG}.log 2>&1
for K in ${KERNELS:-fast_cells resize_level describe_kernel match_pair_kernel quadtree_kernel prep_map}; do
  ncu --set full --clock-control none --import-source on -k regex:$K -s ${SKIP:-2} -c 1 -f -o gpurun_out/prof_${TAG}_$K $CMD > gpurun_out/prof_${TAG}_$K.log 2>&1
done
ls -la gpurun_out
