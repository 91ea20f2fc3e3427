# A/B of the fused pair kernel: parity subset, sanitizer, bench variants, ncu capture
TAG=${1:-r01k}
O=gpurun_out
mkdir -p $O
( time timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "pair or vec_equals_generic or mhd_blast_3d or vec_kernels or implicit_euler or runs_to_t_end" ) > $O/${TAG}_pytest.log 2>&1
tail -15 $O/${TAG}_pytest.log
cat > /tmp/small.py <<'PY'
import numpy as np
from astronomix_b200 import problems as pr, stepping as ti, _lib
s = pr.smooth_random(3, 24, mhd=True, seed=17, dtype=np.float32)
c, p, rv = pr.make(3, 24, mhd=True)
out = ti.integrate_steps(s, c, p, rv, 2)
print("ok", np.isfinite(out).all(), ti.last_info)
PY
PYTHONPATH=. timeout 600 compute-sanitizer --tool memcheck python /tmp/small.py > $O/${TAG}_memcheck.log 2>&1; tail -4 $O/${TAG}_memcheck.log
PYTHONPATH=. timeout 600 compute-sanitizer --tool racecheck python /tmp/small.py > $O/${TAG}_racecheck.log 2>&1; tail -4 $O/${TAG}_racecheck.log
run() { echo "== $@"; env "$@" timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu 2>>$O/${TAG}_bench.err | tee -a $O/${TAG}_bench_variants.jsonl | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], {k:(round(v.get('avg_ms',0),4), v['launches']) for k,v in d['kernels'].items()})"; }
run AMX_PAIR_LD=0
run AMX_PAIR_LD=1
run AMX_NO_PAIR=1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'mag_pair_vec' -s 4 -c 2 \
  -f -o $O/${TAG}_pair python bench.py --steps 2 --warmup 1 --no-cpu > $O/${TAG}_ncu_pair.log 2>&1
tail -2 $O/${TAG}_ncu_pair.log
