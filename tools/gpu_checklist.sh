#!/bin/bash
# One `gpurun` call that (re)validates everything on a B200 box and brings the evidence back in gpurun_out/:
#   /usr/local/graft/bin/gpurun --timeout 1500 -- 'bash tools/gpu_checklist.sh'            # 1 GPU
#   /usr/local/graft/bin/gpurun --gpus 2 --timeout 1500 -- 'bash tools/gpu_checklist.sh 2' # + Z-slab tests / bench on 2 GPUs
N=${1:-1}
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/nvidia_smi.csv 2>&1
timeout 900 python -m pytest tests -q -m gpu -x > gpurun_out/pytest_gpu.log 2>&1
echo "exit $?" >> gpurun_out/pytest_gpu.log
timeout 600 python bench.py > gpurun_out/bench_1gpu.json 2> gpurun_out/bench_1gpu.err
timeout 600 python bench.py --impl reference --steps 10 --warmup 3 > gpurun_out/bench_reference.json 2> gpurun_out/bench_reference.err
if [ "$N" -gt 1 ]; then
    timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node "$N" --master-addr 127.0.0.1 --master-port 29511 \
        bench.py --gpus "$N" > "gpurun_out/bench_${N}gpu.json" 2> "gpurun_out/bench_${N}gpu.err"
fi
tail -3 gpurun_out/pytest_gpu.log
head -c 600 gpurun_out/bench_1gpu.json
