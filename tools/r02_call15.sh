#!/bin/bash
# Round-2 GPU call (2 x B200): slabs balanced by work — fused ghost reads vs NCCL loops
O=gpurun_out/r02c15
mkdir -p $O
for mode in fused nccl; do
    unset SPH_B200_NO_PEER SPH_B200_PEER_FUSED
    [ $mode = nccl ] && export SPH_B200_NO_PEER=1
    timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 \
        bench.py --gpus 2 --steps 20 --warmup 5 > $O/bench_2gpu_$mode.json 2> $O/bench_2gpu_$mode.err; echo "bench2 $mode rc=$?"; tail -4 $O/bench_2gpu_$mode.err | cut -c1-300
    python - $O/bench_2gpu_$mode.json <<'PY'
import json, sys
try:
    d = json.load(open(sys.argv[1]))
    it = d["stats"]["mean_iterations"]
    print("value", d["value"] / 1e6, "M  ms/step", d["ms_per_step"], "it", it, "ms/iteration", d["ms_per_step"] / (it["dfsph_density"] + it["dfsph_divergence"]), "e2e", d["e2e"]["value"] / 1e6)
    print("slab_parity ok", d.get("slab_parity", {}).get("ok"), "pos err", d.get("slab_parity", {}).get("max_rel_position_error"), "slab", d["stats"]["slab"])
    print("top", [(k["name"], round(k["ms_per_launch"] * 1e3, 1), round(k["share"], 3)) for k in d["roofline"]["kernels"][:7]])
except Exception as e:
    print("failed", e)
PY
done
