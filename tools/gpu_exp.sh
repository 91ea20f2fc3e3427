TAG=${1:-r01n}
O=gpurun_out
mkdir -p $O
run() { echo "== $@"; env "$@" timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu 2>>$O/${TAG}_bench.err | tee -a $O/${TAG}_bench_variants.jsonl | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], d['ms_per_step_profiled_pass'], {k:(round(v.get('avg_ms',0),4), v['launches']) for k,v in d['kernels'].items()})"; }
run AMX_X=0
run AMX_LIB_VARIANT=c7b2
run AMX_LIB_VARIANT=c7b3
run AMX_NO_PAIR=1
run AMX_NO_PAIR=1 AMX_LIB_VARIANT=m8b2 AMX_VEC_CW=8
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "vec_hydro or mhd_blast_3d or kelvin or conservation or vec_kernels" 2>&1 | tail -4
