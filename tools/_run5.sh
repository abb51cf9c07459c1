mkdir -p gpurun_out
cd tools/ubench && nvcc -gencode arch=compute_100a,code=sm_100a -o tma_test tma_test.cu -lcuda 2>&1 | tail -3
( for m in 0 1 2 3 4; do for x in "4 16" "3 16" "4 12" "8 32"; do echo "== mode $m x4,x8 = $x"; timeout 30 ./tma_test $m $x 2>&1 | grep -E "^kernel|mismatch" ; done; done ) > ../../gpurun_out/tma_test2.txt 2>&1
cd ../..
python bench.py --steps 8 --warmup 3 --no-cpu-baseline --no-bruteforce > gpurun_out/b5_default.json 2> gpurun_out/b5_default.err
for v in w4 w2; do
  DSX_LIB=$PWD/diasss_b200/variants/libdiasss_b200_$v.so python -m pytest tests/test_gpu_extract.py -m gpu -x -q 2>&1 | tail -2
  DSX_LIB=$PWD/diasss_b200/variants/libdiasss_b200_$v.so python bench.py --steps 8 --warmup 3 --no-cpu-baseline --no-bruteforce --no-e2e > gpurun_out/b5_$v.json 2> gpurun_out/b5_$v.err
done
