# quick correctness + kernel timing after a kernel change; usage: bash tools/gpu_quick.sh TAG [pytest -k expr]
TAG=${1:-rXX}
K=${2:-"vec_hydro or vec_kernels_fp32 or mhd_blast_3d or conservation or pair_blast"}
mkdir -p gpurun_out
( timeout 500 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "$K" ) > gpurun_out/${TAG}_pytest_quick.log 2>&1
tail -3 gpurun_out/${TAG}_pytest_quick.log
timeout 120 python tools/kernel_bench.py 2>&1 | tail -1 | tee -a gpurun_out/${TAG}_kbench.jsonl
