# A/B of kernel variants: parity subset on the default library, then bench.py per variant library.
# usage: gpurun --timeout 900 -- 'bash tools/gpu_ab.sh r04b "" alu'
TAG=${1:-rXX}; shift
O=gpurun_out; mkdir -p $O
( time timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "vec_ or mhd_blast_3d or conservation or pair_blast or slab_driver_single or tiled_fp32 or translation" ) > $O/${TAG}_pytest_subset.log 2>&1
tail -4 $O/${TAG}_pytest_subset.log
for v in "$@"; do
  name=${v:-default}
  AMX_LIB_VARIANT=$v timeout 200 python bench.py --steps 20 --warmup 3 --no-cpu > $O/${TAG}_bench_${name}.json 2> $O/${TAG}_bench_${name}.err
  python - <<PY
import json
d=json.load(open("$O/${TAG}_bench_${name}.json"))
print("$name", d["value"], d["ms_per_step"], {k:(round(v.get("avg_ms",0),4), v["launches"]) for k,v in d["kernels"].items()})
PY
done
