set -x
python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "vec_kernels or (vec_equals and (3-32 or 2-512) and midpoint)" 2>&1 | tail -5 > gpurun_out/r01i_pytest_vec.log
cat gpurun_out/r01i_pytest_vec.log
python bench.py --steps 20 --warmup 3 --no-cpu > gpurun_out/r01i_bench.json 2> gpurun_out/r01i_bench.err
cat gpurun_out/r01i_bench.json
tail -3 gpurun_out/r01i_bench.err
