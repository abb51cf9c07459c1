#!/bin/bash
# Usage (under gpurun --gpus N): bash tools/scale_round.sh <N> <tag> [quick]
# One multi-GPU measurement session: concurrent H2D bandwidth, the default survey at N (and N/2) GPUs, the
# high-density sweep (BASELINE config 5) at N GPUs.  Every JSON line lands in gpurun_out/ with its clocks.
N=${1:-8}; TAG=${2:-r02}; QUICK=${3:-}
mkdir -p gpurun_out
run() { python -m torch.distributed.run --nnodes=1 --nproc-per-node $1 --master-addr 127.0.0.1 --master-port $2 "${@:3}"; }
run $N 29601 tools/h2d_concurrent.py > gpurun_out/h2d_${TAG}_n$N.json 2> gpurun_out/h2d_${TAG}_n$N.err
run $N 29602 tools/h2d_concurrent.py --bind-numa > gpurun_out/h2d_${TAG}_n${N}_numa.json 2>> gpurun_out/h2d_${TAG}_n$N.err
nvidia-smi topo -m > gpurun_out/topo_${TAG}_n$N.txt 2>&1
S="--steps 20 --warmup 5"; [ -n "$QUICK" ] && S="--steps 4 --warmup 3"
run $N 29603 bench.py --gpus $N $S --no-bruteforce > gpurun_out/bench_${TAG}_n$N.json 2> gpurun_out/bench_${TAG}_n$N.err
if [ -n "$HALF" ] && [ $N -gt 2 ]; then
  H=$((N/2))
  run $H 29604 bench.py --gpus $H $S --no-bruteforce > gpurun_out/bench_${TAG}_n$H.json 2> gpurun_out/bench_${TAG}_n$H.err
fi
for NF in 5000 20000 50000; do
  [ -n "$QUICK" ] && [ $NF != 5000 ] && continue
  run $N 29605 bench.py --gpus $N --steps 5 --warmup 3 --no-bruteforce --nfeatures $NF > gpurun_out/bench_${TAG}_density_${NF}_n$N.json 2> gpurun_out/bench_${TAG}_density_${NF}_n$N.err
done
tail -c 400 gpurun_out/bench_${TAG}_*.err
