set -x
python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "vec_hydro_fp32 or vec_kernels or (vec_hydro_stage and 3-32)" 2>&1 | tail -5 > gpurun_out/r01g_pytest_vec.log
cat gpurun_out/r01g_pytest_vec.log
python bench.py --steps 20 --warmup 3 --no-cpu > gpurun_out/r01g_bench.json 2> gpurun_out/r01g_bench.err
cat gpurun_out/r01g_bench.json
tail -3 gpurun_out/r01g_bench.err
ncu --set full --clock-control none --import-source on -k regex:hydro_stage_vec -s 8 -c 2 -o gpurun_out/r01g_stage python bench.py --steps 4 --warmup 3 --no-cpu > gpurun_out/r01g_ncu_stage.log 2>&1
