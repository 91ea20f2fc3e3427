# A/B of the fused pair kernel: parity subset, bench variants, ncu capture
TAG=${1:-r01k}
O=gpurun_out
mkdir -p $O
( time timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "pair or van_albada_pp or (vec_equals_generic and 3-32) or mhd_blast_3d" ) > $O/${TAG}_pytest.log 2>&1
tail -15 $O/${TAG}_pytest.log
run() { echo "== $@"; env "$@" timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu 2>>$O/${TAG}_bench.err | tee -a $O/${TAG}_bench_variants.jsonl | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], {k:(round(v.get('avg_ms',0),4), v['launches']) for k,v in d['kernels'].items()})"; }
run AMX_PAIR_LD=0
run AMX_PAIR_LD=1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'mag_pair_vec' -s 4 -c 1 \
  -f -o $O/${TAG}_pair python bench.py --steps 2 --warmup 1 --no-cpu > $O/${TAG}_ncu_pair.log 2>&1
tail -2 $O/${TAG}_ncu_pair.log
