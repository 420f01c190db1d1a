#!/bin/bash
# Round-2 GPU call: bench.py (both arms) on the headline config, bench lines of the other BASELINE configs
O=gpurun_out/r02c9
mkdir -p $O
nproc > $O/nproc.txt; lscpu | grep "Model name" >> $O/nproc.txt
timeout 900 python bench.py --steps 20 --warmup 5 > $O/bench_c2p_dfsph.json 2> $O/bench_c2p_dfsph.err; echo "bench rc=$?"; tail -3 $O/bench_c2p_dfsph.err
python - $O/bench_c2p_dfsph.json <<'PY'
import json, sys
d = json.load(open(sys.argv[1]))
print("value", d["value"] / 1e6, "M  ms/step", d["ms_per_step"], "e2e", d["e2e"]["value"] / 1e6, "launches", d["gpu_launches"])
print("stats", json.dumps(d["stats"])[:600])
print("cpu", json.dumps(d["cpu_baseline"])[:900])
r = d["roofline"]; print("roofline", r["kernel"], r["achieved"], r["frac"], r["traffic"], "density:", json.dumps(r.get("density_kernel"))[:300])
PY
timeout 1500 python bench.py --impl reference --steps 10 --warmup 3 > $O/bench_reference.json 2> $O/bench_reference.err; echo "reference rc=$?"; tail -3 $O/bench_reference.err; head -c 1500 $O/bench_reference.json; echo
for c in c2_wcsph c3_bath c4_buckling; do
    timeout 600 python bench.py --config $c --steps 20 --warmup 5 > $O/bench_$c.json 2> $O/bench_$c.err; echo "$c rc=$?"; tail -2 $O/bench_$c.err
    python - $O/bench_$c.json <<'PY'
import json, sys
try:
    d = json.load(open(sys.argv[1]))
    print("  value", d["value"] / 1e6, "M  ms/step", d["ms_per_step"], "e2e", d["e2e"]["value"] / 1e6, "it", d["stats"]["mean_iterations"])
    print("  cpu", json.dumps(d["cpu_baseline"])[:700])
    print("  top", [(k["name"], round(k["ms_per_launch"] * 1e3, 1), round(k["share"], 3)) for k in d["roofline"]["kernels"][:5]])
except Exception as e:
    print("  failed", e)
PY
done
