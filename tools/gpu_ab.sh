run() { echo "$@"; env "$@" python bench.py --steps 10 --warmup 3 --no-cpu | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], {k:round(v['ms_total']/v.get('evaluations',v['launches']),4) for k,v in d['kernels'].items()})"; }
run AMX_VEC_PF=0
run AMX_VEC_PF=1
run AMX_VEC_PF=2
run AMX_VEC_PF=4
