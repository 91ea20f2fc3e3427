TAG=${1:-r01o}
O=gpurun_out
mkdir -p $O
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29531 tools/comm_micro.py > $O/${TAG}_comm_micro.log 2>&1
grep -E "us per call|Error|slab\]" $O/${TAG}_comm_micro.log | head -20
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 tools/slab_check.py > $O/${TAG}_slab_check.log 2>&1
grep -E "SLAB_CHECK|bitwise|slab\]|Error|error" $O/${TAG}_slab_check.log | tail -8
for P2P in 1 0; do
AMX_SLAB_P2P=$P2P timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 2951$P2P bench.py --gpus 2 --steps 20 --warmup 3 > $O/${TAG}_bench_n2_p2p$P2P.json 2> $O/${TAG}_bench_n2_p2p$P2P.err
grep -E "slab\]|Error" $O/${TAG}_bench_n2_p2p$P2P.err | head -3
python -c "
import json; d=json.loads(open('$O/${TAG}_bench_n2_p2p$P2P.json').read().strip().splitlines()[-1]); print('p2p$P2P', d['value'], d['ms_per_step'], d['e2e']['value'], d['gpu_launches'])"
done
