# bench.py per variant library (AMX_LIB_VARIANT); usage: bash tools/gpu_bench_variants.sh TAG v1 v2 ...
TAG=${1:-rXX}; shift
O=gpurun_out; mkdir -p $O
for v in "$@"; do
  name=${v:-default}
  AMX_LIB_VARIANT=$v timeout 200 python bench.py --steps 20 --warmup 3 --no-cpu > $O/${TAG}_bench_${name}.json 2> $O/${TAG}_bench_${name}.err || tail -3 $O/${TAG}_bench_${name}.err
  python - <<PY
import json
try:
    d=json.load(open("$O/${TAG}_bench_${name}.json"))
    print("$name", "%.4g" % d["value"], round(d["ms_per_step"],4), {k:(round(v.get("avg_ms",0),4), v["launches"]) for k,v in d["kernels"].items() if v.get("avg_ms")})
except Exception as e: print("$name failed", e)
PY
done
