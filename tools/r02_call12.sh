#!/bin/bash
# Round-2 GPU call (2 x B200): peer-memory solver loops vs the NCCL loops: slab parity tests + bench lines of both
O=gpurun_out/r02c12
mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_slab.py -q -m gpu > $O/pytest_slab_peer.log 2>&1; echo "exit $?" >> $O/pytest_slab_peer.log; tail -12 $O/pytest_slab_peer.log | cut -c1-400
for mode in peer nccl; do
    if [ $mode = nccl ]; then export SPH_B200_NO_PEER=1; else unset SPH_B200_NO_PEER; fi
    timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 \
        bench.py --gpus 2 --steps 20 --warmup 5 > $O/bench_2gpu_$mode.json 2> $O/bench_2gpu_$mode.err; echo "bench2 $mode rc=$?"; tail -4 $O/bench_2gpu_$mode.err | cut -c1-300
    python - $O/bench_2gpu_$mode.json <<'PY'
import json, sys
try:
    d = json.load(open(sys.argv[1]))
    it = d["stats"]["mean_iterations"]
    print("value", d["value"] / 1e6, "M  ms/step", d["ms_per_step"], "it", it, "ms/iteration", d["ms_per_step"] / (it["dfsph_density"] + it["dfsph_divergence"]))
    print("slab_parity", d.get("slab_parity"))
    print("top", [(k["name"], round(k["ms_per_launch"] * 1e3, 1), round(k["share"], 3)) for k in d["roofline"]["kernels"][:8]])
except Exception as e:
    print("failed", e)
PY
done
