set -x
python -m pytest tests -m gpu -x -q 2>&1 | tail -15 > gpurun_out/r01b_pytest.log
python bench.py --steps 20 --warmup 3 > gpurun_out/r01b_bench.json 2> gpurun_out/r01b_bench.err
cat gpurun_out/r01b_bench.json
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r01b_launches.csv python bench.py --steps 4 --warmup 3 --no-cpu > gpurun_out/r01b_ncu_bench.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:hydro_stage_tiled -s 8 -c 2 -o gpurun_out/r01b_stage python bench.py --steps 4 --warmup 3 --no-cpu > gpurun_out/r01b_ncu_stage.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:mag_iter_tiled -s 8 -c 2 -o gpurun_out/r01b_mag python bench.py --steps 4 --warmup 3 --no-cpu > gpurun_out/r01b_ncu_mag.log 2>&1
ls -la gpurun_out
