#!/bin/bash
# Usage (under gpurun): bash tools/traffic_round.sh <tag>
# DRAM bytes, duration and pipe utilisation of every launch of one bench step (few metrics -> few replays), as CSV.
TAG=${1:-r01}
mkdir -p gpurun_out
KREGEX='regex:(resize_level|fast_cells|quadtree_kernel|describe_kernel|finalize_kernel|georef_kernel|match_pair_kernel|match_prepare|scc_merge|scan_counts|emit_rows)'
M=dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active,sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active,sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active,smsp__issue_active.avg.pct_of_peak_sustained_active
ncu --metrics $M --clock-control none -k "$KREGEX" -c 400 --csv --log-file gpurun_out/traffic_${TAG}.csv \
    python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu-baseline --no-bruteforce > gpurun_out/traffic_${TAG}.log 2>&1
