#!/bin/bash
# Round-2 GPU call: stored pair weights — bitwise list/walk check, all GPU tests, per-kernel times, ncu of the iteration sweeps
O=gpurun_out/r02c13
mkdir -p $O
timeout 300 python tools/debug_lists.py small > $O/debug_small.log 2>&1; echo "debug small rc=$?"; grep -v "differing rows        0" $O/debug_small.log | grep -v "^Dimension\|^grid size\|^Number of\|^Fluid particle\|^No rigid" | cut -c1-300 | head -20
timeout 900 python -m pytest tests -q -m gpu > $O/pytest_gpu.log 2>&1
echo "exit $?" >> $O/pytest_gpu.log
grep -E "passed|failed|FAILED|exit" $O/pytest_gpu.log | head -20
for sc in dam_break_1m_dfsph:1000 dam_break_1m_wcsph:300; do
    name=${sc%%:*}; settle=${sc##*:}
    timeout 600 python tools/scene_step.py --scene data/scenes/$name.json --settle $settle --steps 20 > $O/step_$name.json 2> $O/step_$name.err
    python - $O/step_$name.json $name <<'PY'
import json, sys
try:
    d = json.load(open(sys.argv[1]))
    print(sys.argv[2], "ms/step", round(d["ms_per_step"], 3), "iters", d["stats"]["total_dfsph_iterations"], "kernel ms", round(d["kernel_ms_per_step"], 3))
    for k in d["kernels"][:7]: print(f"  {k['name']:48s} {k['launches_per_step']:6.2f} x {k['ms_per_launch']*1e3:8.1f} us  {k['share']:.3f}")
except Exception as e:
    print(sys.argv[2], "failed", e)
PY
done
NCU="ncu --profile-from-start off --set full --import-source on --clock-control none"
timeout 900 $NCU -k regex:'kb_dfsph_correct|kb_dfsph_density_change' -c 2 -f -o $O/ncu_brick_iter \
    python tools/scene_step.py --scene data/scenes/dam_break_1m_dfsph.json --settle 1000 --steps 1 --cuda-profiler --no-profile-pass > $O/ncu_brick_iter.log 2>&1
[ -f $O/ncu_brick_iter.ncu-rep ] && ncu -i $O/ncu_brick_iter.ncu-rep --page raw --csv > $O/ncu_brick_iter_raw.csv 2>/dev/null
python profiles/tools/ncu_summary.py $O/ncu_brick_iter_raw.csv
