#!/bin/bash
# Round-2 GPU call (2 x B200): Z-slab parity tests and the 2-GPU bench line with the brick kernels
O=gpurun_out/r02c10
mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_slab.py -q -m gpu > $O/pytest_slab.log 2>&1; echo "exit $?" >> $O/pytest_slab.log; tail -15 $O/pytest_slab.log | cut -c1-300
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 \
    bench.py --gpus 2 --steps 20 --warmup 5 > $O/bench_2gpu.json 2> $O/bench_2gpu.err; echo "bench2 rc=$?"; tail -5 $O/bench_2gpu.err | cut -c1-300
python - $O/bench_2gpu.json <<'PY'
import json, sys
try:
    d = json.load(open(sys.argv[1]))
    print("value", d["value"] / 1e6, "M  ms/step", d["ms_per_step"], "e2e", d["e2e"]["value"] / 1e6, "it", d["stats"]["mean_iterations"])
    print("slab_parity", d.get("slab_parity"))
    print("slab", d["stats"].get("slab"))
    print("top", [(k["name"], round(k["ms_per_launch"] * 1e3, 1), round(k["share"], 3)) for k in d["roofline"]["kernels"][:6]])
except Exception as e:
    print("failed", e)
PY
