set -x
python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "vec" 2>&1 | tail -15 > gpurun_out/r01e_pytest_vec.log
cat gpurun_out/r01e_pytest_vec.log
python bench.py --steps 20 --warmup 3 --no-cpu > gpurun_out/r01e_bench.json 2> gpurun_out/r01e_bench.err
cat gpurun_out/r01e_bench.json
ncu --set full --clock-control none --import-source on -k regex:mag_iter_vec -s 8 -c 2 -o gpurun_out/r01e_mag python bench.py --steps 4 --warmup 3 --no-cpu > gpurun_out/r01e_ncu_mag.log 2>&1
