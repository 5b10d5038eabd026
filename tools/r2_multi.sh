#!/bin/bash
# multi-GPU session: N = $1 GPUs of one box.  Equality test (2 ranks), the pipeline step with the descriptor all-gather
# inside (c2), and the retrieval split (c4).
N=${1:-2}
mkdir -p gpurun_out
T0=$(date +%s)
stamp() { echo "[$(( $(date +%s) - T0 )) s] $*" | tee -a gpurun_out/r2_multi${N}_steps.log; }
: > gpurun_out/r2_multi${N}_steps.log
RUN="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511"
nvidia-smi -L | head -8 >> gpurun_out/r2_multi${N}_steps.log
timeout 300 python -m pytest tests/test_dist_gpu.py -m gpu -q > gpurun_out/r2_multi${N}_test.log 2>&1
stamp "test_dist_gpu: $(tail -1 gpurun_out/r2_multi${N}_test.log)"
NCCL_DEBUG=WARN timeout 400 $RUN bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/r2_multi${N}_c2.log 2>&1
stamp "c2 x$N: $(grep -o '"value": [0-9.]*' gpurun_out/r2_multi${N}_c2.log | head -1)"
timeout 400 $RUN bench.py --gpus $N --workload c4 --steps 5 --warmup 2 > gpurun_out/r2_multi${N}_c4.log 2>&1
stamp "c4 x$N: $(grep -o '"ms_per_step": [0-9.]*' gpurun_out/r2_multi${N}_c4.log | head -1)"
cat gpurun_out/r2_multi${N}_steps.log
