TAG=${1:-r01n}
O=gpurun_out
mkdir -p $O
run() { echo "== $@"; env "$@" timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu 2>>$O/${TAG}_bench.err | tee -a $O/${TAG}_bench_variants.jsonl | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], d['ms_per_step_profiled_pass'], {k:(round(v.get('avg_ms',0),4), v['launches']) for k,v in d['kernels'].items()})"; }
run AMX_PAIR_PF=0
run AMX_PAIR_PF=1
run AMX_PAIR_PF=2
run AMX_PAIR_PF=3
run AMX_PAIR_PF=4
