mkdir -p gpurun_out
python -m pytest tests/test_gpu_extract.py tests/test_gpu_configs.py -m gpu -x -q 2>&1 | tail -3
python bench.py --steps 8 --warmup 3 --no-cpu-baseline --no-bruteforce > gpurun_out/b6_default.json 2> gpurun_out/b6_default.err
DSX_LIB=$PWD/diasss_b200/variants/libdiasss_b200_prof.so python tools/fast_phase_profile.py > gpurun_out/fast_phase_profile.json 2> gpurun_out/fast_phase_profile.err
