#!/bin/bash
# Round-2 GPU call (8 x B200): weak-scaling line of the headline config and BASELINE config 5 (10 M dam break)
O=gpurun_out/r02c23
mkdir -p $O
nvidia-smi topo -m > $O/topo.txt 2>&1
run() {  # name, extra args
    timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 \
        bench.py --gpus 8 --steps 20 --warmup 5 $2 > $O/bench_8gpu_$1.json 2> $O/bench_8gpu_$1.err; echo "bench8 $1 rc=$?"; tail -4 $O/bench_8gpu_$1.err | cut -c1-300
    python - $O/bench_8gpu_$1.json <<'PY'
import json, sys
try:
    d = json.load(open(sys.argv[1]))
    it = d["stats"]["mean_iterations"]
    print("value", d["value"] / 1e6, "M  ms/step", d["ms_per_step"], "it", it, "ms/iteration", d["ms_per_step"] / (it["dfsph_density"] + it["dfsph_divergence"]), "e2e", d["e2e"]["value"] / 1e6)
    print("slab_parity", d.get("slab_parity"))
    print("slab", d["stats"]["slab"], "bricks", d["stats"]["bricks"])
    print("top", [(k["name"], round(k["ms_per_launch"] * 1e3, 1), round(k["share"], 3)) for k in d["roofline"]["kernels"][:8]])
except Exception as e:
    print("failed", e)
PY
}
run c2p ""
run c5_pressurised "--config c5_dam10m --settle 1500"
