# N-GPU check + bench; usage: gpurun --gpus N --timeout 900 -- 'bash tools/gpu_n.sh N TAG'
N=${1:-2}; TAG=${2:-rXX}
O=gpurun_out
mkdir -p $O
AMX_CHECK_N=${AMX_CHECK_N:-128} timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 tools/slab_check.py > $O/${TAG}_slab_check_n$N.log 2>&1
grep -E "SLAB_CHECK|bitwise|slab\]|Error|error" $O/${TAG}_slab_check_n$N.log | tail -12
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29518 bench.py --gpus $N --steps 20 --warmup 5 > $O/${TAG}_bench_n$N.json 2> $O/${TAG}_bench_n$N.err
grep -E "slab\]|Error|error" $O/${TAG}_bench_n$N.err | head -5
python -c "
import json; d=json.loads(open('$O/${TAG}_bench_n$N.json').read().strip().splitlines()[-1]); print('N=$N', d['value'], d['ms_per_step'], d['e2e']['value'], d['gpu_launches'], json.dumps(d['config'])[:1500])"
