TAG=${1:-r01o}
O=gpurun_out
mkdir -p $O
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 tools/slab_check.py > $O/${TAG}_slab_check.log 2>&1
grep -E "SLAB_CHECK|bitwise" $O/${TAG}_slab_check.log | tail -6
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29518 bench.py --gpus 2 --steps 10 --warmup 3 > $O/${TAG}_bench_n2.json 2> $O/${TAG}_bench_n2.err
tail -1 $O/${TAG}_bench_n2.json; tail -3 $O/${TAG}_bench_n2.err
