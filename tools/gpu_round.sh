# Full GPU round: [parity suite,] smoke, bench (native + reference arm), ncu launch list, ncu full captures.
# usage: gpurun --timeout 1500 -- 'bash tools/gpu_round.sh r01j [nopytest]'
TAG=${1:-rXX}
O=gpurun_out
mkdir -p $O
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $O/${TAG}_smi.txt
if [ "$2" != "nopytest" ]; then
( time timeout 1100 python -m pytest tests -m gpu -x -q ) > $O/${TAG}_pytest.log 2>&1
tail -5 $O/${TAG}_pytest.log
fi
timeout 300 python __graft_entry__.py smoke > $O/${TAG}_smoke.log 2>&1; tail -2 $O/${TAG}_smoke.log
timeout 600 python bench.py --steps 20 --warmup 5 > $O/${TAG}_bench.json 2> $O/${TAG}_bench.err
cat $O/${TAG}_bench.json; tail -3 $O/${TAG}_bench.err
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > $O/${TAG}_bench_ref.json 2> $O/${TAG}_bench_ref.err
cat $O/${TAG}_bench_ref.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/${TAG}_launches.csv \
  python bench.py --steps 2 --warmup 3 --no-cpu --no-secondary > $O/${TAG}_ncu_launch.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'hydro_stage_vec|mag_pair_vec|mag_final' -s 8 -c 5 \
  -f -o $O/${TAG}_full python bench.py --steps 2 --warmup 3 --no-cpu --no-secondary > $O/${TAG}_ncu_full.log 2>&1
tail -2 $O/${TAG}_ncu_full.log
ls -la $O | tail -12
