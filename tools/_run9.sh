mkdir -p gpurun_out
python -m pytest tests/test_gpu_extract.py tests/test_gpu_configs.py tests/test_gpu_match.py -m gpu -x -q 2>&1 | tail -3
DSX_DESCRIBE_DENSE=1 python -m pytest tests/test_gpu_extract.py -m gpu -x -q 2>&1 | tail -3
python bench.py --steps 8 --warmup 3 --no-cpu-baseline --no-bruteforce > gpurun_out/b9.json 2> gpurun_out/b9.err
for t in 2 1 0; do
DSX_FAST_TMA=$t DSX_LIB=$PWD/diasss_b200/variants/libdiasss_b200_prof.so python tools/fast_phase_profile.py > gpurun_out/fast_phase_profile_tma$t.json 2> gpurun_out/fast_phase_profile.err
done
DSX_FAST_TMA=0 python bench.py --steps 8 --warmup 3 --no-cpu-baseline --no-bruteforce --no-e2e > gpurun_out/b9_tma0.json 2> gpurun_out/b9_tma0.err
python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-bruteforce --no-e2e --nfeatures 20000 > gpurun_out/b9_nf20k_dense.json 2> gpurun_out/b9_nf20k_dense.err
DSX_DESCRIBE_DENSE=0 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-bruteforce --no-e2e --nfeatures 20000 > gpurun_out/b9_nf20k_window.json 2> gpurun_out/b9_nf20k_window.err
DSX_SLOW_TESTS=1 python -m pytest "tests/test_gpu_match.py::test_high_density_pair_vs_oracle" -m gpu -x -q 2>&1 | tail -3 > gpurun_out/pytest_highdensity_50k.log
