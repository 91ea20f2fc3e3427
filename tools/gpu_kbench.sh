# kernel_bench.py per variant library; usage: bash tools/gpu_kbench.sh TAG "what" v1 v2 ...   ("" = default library)
TAG=${1:-rXX}; WHAT=${2:-stage,stage2,iter,pair}; shift; shift
mkdir -p gpurun_out
for v in "$@"; do
  AMX_LIB_VARIANT=$v timeout 120 python tools/kernel_bench.py --what $WHAT 2>&1 | tail -1 | tee -a gpurun_out/${TAG}_kbench.jsonl
done
