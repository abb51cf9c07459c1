run() { python -m torch.distributed.run --nnodes=1 --nproc-per-node $1 --master-addr 127.0.0.1 --master-port $2 "${@:3}"; }
run 8 29621 bench.py --gpus 8 --steps 20 --warmup 5 --no-bruteforce > gpurun_out/bench_m_n8.json 2> gpurun_out/bench_m_n8.err
run 8 29622 bench.py --gpus 8 --steps 10 --warmup 3 --no-bruteforce --no-h2d-split > gpurun_out/bench_m_n8_equal.json 2> gpurun_out/bench_m_n8_equal.err
run 8 29623 bench.py --gpus 8 --steps 5 --warmup 3 --no-bruteforce --nfeatures 20000 > gpurun_out/bench_m_density_20000_n8.json 2> gpurun_out/bench_m_density_20000_n8.err
for f in bench_m_n8 bench_m_n8_equal bench_m_density_20000_n8; do python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/$f.json').read().strip().splitlines()[-1])
    print('$f', d['ms_per_step'], d['e2e']['ms_per_step'], d['e2e']['isolated_ms_per_step'], d['e2e'].get('image_split'), d['e2e']['rows6_sha256'], d['config']['output_sha256']['rows6'], d['clocks'])
except Exception as e: print('$f', 'ERR', e)
PY
done
tail -c 300 gpurun_out/bench_m_n8.err
