mkdir -p gpurun_out
python -m pytest tests/test_gpu_extract.py tests/test_gpu_configs.py -m gpu -x -q 2>&1 | tail -3
python bench.py --steps 8 --warmup 3 --no-cpu-baseline --no-bruteforce > gpurun_out/b7_tma2.json 2> gpurun_out/b7_tma2.err
DSX_FAST_TMA=1 python bench.py --steps 8 --warmup 3 --no-cpu-baseline --no-bruteforce --no-e2e > gpurun_out/b7_tma1.json 2> gpurun_out/b7_tma1.err
DSX_LIB=$PWD/diasss_b200/variants/libdiasss_b200_prof.so python tools/fast_phase_profile.py > gpurun_out/fast_phase_profile_tma2.json 2> gpurun_out/fast_phase_profile.err
DSX_FAST_TMA=1 DSX_LIB=$PWD/diasss_b200/variants/libdiasss_b200_prof.so python tools/fast_phase_profile.py > gpurun_out/fast_phase_profile_tma1.json 2>> gpurun_out/fast_phase_profile.err
