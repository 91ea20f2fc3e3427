TAG=${1:-r01n}
O=gpurun_out
mkdir -p $O
run() { echo "== $@"; env "$@" timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu 2>>$O/${TAG}_bench.err | tee -a $O/${TAG}_bench_variants.jsonl | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], d['ms_per_step_profiled_pass'], {k:(round(v.get('avg_ms',0),4), v['launches']) for k,v in d['kernels'].items()})"; }
run AMX_X=0
run AMX_LIB_VARIANT=mm3
AMX_LIB_VARIANT=mm3 timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "mhd_blast_3d or vec_hydro_fp32 or kelvin or conservation" 2>&1 | tail -4
