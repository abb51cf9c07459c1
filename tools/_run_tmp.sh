python -m pytest tests/test_gpu_match.py tests/test_gpu_configs.py tests/test_gpu_ref.py tests/test_gpu_peer.py tests/test_gpu_demo.py -m gpu -x -q 2>&1 | tail -3
DSX_SLOW_TESTS=1 python -m pytest tests/test_gpu_match.py -m gpu -x -q -k "high_density" 2>&1 | tail -3
B="python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-e2e"
$B 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('new', d['ms_per_step'], d['stages_ms_per_step']['match'], d['config']['output_sha256'])"
DSX_MATCH_COMPACT=0 $B 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('noncompact', d['ms_per_step'], d['stages_ms_per_step']['match'], d['config']['output_sha256']['rows6'])"
$B --nfeatures 20000 --steps 3 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('20k', d['ms_per_step'], d['stages_ms_per_step'], d['config']['output_sha256']['rows6'])"
