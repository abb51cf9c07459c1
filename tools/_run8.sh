mkdir -p gpurun_out
python -m pytest tests/test_gpu_extract.py tests/test_gpu_configs.py -m gpu -x -q 2>&1 | tail -3
python bench.py --steps 8 --warmup 3 --no-cpu-baseline --no-bruteforce > gpurun_out/b8.json 2> gpurun_out/b8.err
for t in 2 1 0; do
DSX_FAST_TMA=$t DSX_LIB=$PWD/diasss_b200/variants/libdiasss_b200_prof.so python tools/fast_phase_profile.py > gpurun_out/fast_phase_profile_tma$t.json 2> gpurun_out/fast_phase_profile.err
done
DSX_FAST_TMA=0 python bench.py --steps 8 --warmup 3 --no-cpu-baseline --no-bruteforce --no-e2e > gpurun_out/b8_tma0.json 2> gpurun_out/b8_tma0.err
DSX_SLOW_TESTS=1 python -m pytest "tests/test_gpu_match.py::test_high_density_pair_vs_oracle" -m gpu -x -q 2>&1 | tail -3 > gpurun_out/pytest_highdensity_50k.log
