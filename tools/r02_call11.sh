#!/bin/bash
# Round-2 GPU call: C4 parity anatomy; compute-sanitizer (memcheck + racecheck) on small WCSPH / DFSPH runs
O=gpurun_out/r02c11
mkdir -p $O
timeout 600 python tools/c4_parity.py 12 > $O/c4_parity.log 2> $O/c4_parity.err; echo "c4 rc=$?"; cat $O/c4_parity.log | cut -c1-330; tail -3 $O/c4_parity.err
for sc in dam_break_8k_wcsph; do
  for tool in memcheck racecheck; do
    timeout 900 compute-sanitizer --tool $tool --print-limit 20 python tools/scene_step.py --scene data/scenes/$sc.json --settle 2 --steps 2 --no-profile-pass > $O/sanitizer_${tool}_$sc.log 2>&1
    echo "$tool $sc rc=$?"; grep -E "ERROR SUMMARY|RACECHECK SUMMARY|hazard|Invalid" $O/sanitizer_${tool}_$sc.log | head -8
  done
done
python - <<'PY' > gpurun_out/r02c11/small_dfsph.json
import json
sc = json.load(open("data/scenes/dam_break_8k_wcsph.json"))
sc["Configuration"]["simulationMethod"] = "dfsph"; sc["Configuration"]["timeStepSize"] = 1e-3
sc["FluidBlocks"][0]["velocity"] = [0.0, -1.0, 0.0]
print(json.dumps(sc))
PY
for tool in memcheck racecheck; do
    timeout 900 compute-sanitizer --tool $tool --print-limit 20 python tools/scene_step.py --scene $O/small_dfsph.json --settle 2 --steps 2 --no-profile-pass > $O/sanitizer_${tool}_dfsph_8k.log 2>&1
    echo "$tool dfsph rc=$?"; grep -E "ERROR SUMMARY|RACECHECK SUMMARY|hazard|Invalid" $O/sanitizer_${tool}_dfsph_8k.log | head -8
done
