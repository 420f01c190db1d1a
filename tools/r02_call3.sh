#!/bin/bash
# Round-2 GPU call 3: where the brick-list path departs from the window walk, all GPU tests, per-kernel times, ncu of the brick kernels
O=gpurun_out/r02c3
mkdir -p $O
timeout 300 python tools/debug_lists.py small > $O/debug_small.log 2>&1; echo "debug small rc=$?"; grep -v "differing rows        0" $O/debug_small.log | cut -c1-400 | head -40
timeout 600 python tools/debug_lists.py big > $O/debug_big.log 2>&1; echo "debug big rc=$?"; grep -v "differing rows        0" $O/debug_big.log | cut -c1-400 | head -40
timeout 900 python -m pytest tests -q -m gpu > $O/pytest_gpu.log 2>&1
echo "exit $?" >> $O/pytest_gpu.log
grep -E "passed|failed|FAILED|exit" $O/pytest_gpu.log | head -40
timeout 600 python tools/scene_step.py --scene data/scenes/dam_break_1m_dfsph.json --settle 1000 --steps 20 > $O/step_dfsph_press.json 2> $O/step_dfsph_press.err
python - $O/step_dfsph_press.json <<'PY'
import json, sys
d = json.load(open(sys.argv[1]))
print(d["ms_per_step"], d["stats"]["total_dfsph_iterations"], d["kernel_ms_per_step"])
for k in d["kernels"]: print(f"  {k['name']:48s} {k['launches_per_step']:6.2f} x {k['ms_per_launch']*1e3:8.1f} us  {k['share']:.3f}")
PY
NCU="ncu --profile-from-start off --set full --import-source on --clock-control none"
timeout 900 $NCU -k regex:'kb_dfsph_correct|kb_dfsph_density_change' -c 2 -f -o $O/ncu_brick_iter \
    python tools/scene_step.py --scene data/scenes/dam_break_1m_dfsph.json --settle 1000 --steps 1 --cuda-profiler --no-profile-pass > $O/ncu_brick_iter.log 2>&1
timeout 900 $NCU -k regex:'kb_build' -c 1 -f -o $O/ncu_brick_build \
    python tools/scene_step.py --scene data/scenes/dam_break_1m_dfsph.json --settle 1000 --steps 1 --cuda-profiler --no-profile-pass > $O/ncu_brick_build.log 2>&1
for r in ncu_brick_iter ncu_brick_build; do
    [ -f $O/$r.ncu-rep ] && ncu -i $O/$r.ncu-rep --page raw --csv > $O/${r}_raw.csv 2>/dev/null
done
python profiles/tools/ncu_summary.py $O/ncu_brick_iter_raw.csv $O/ncu_brick_build_raw.csv
