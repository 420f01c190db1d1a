#!/bin/bash
# Round-2 GPU call: sorted-row bricks — correctness, per-kernel times of the default (8x4x4) and the 4x4x4 variants, ncu
O=gpurun_out/r02c5
mkdir -p $O
timeout 300 python tools/debug_lists.py small > $O/debug_small.log 2>&1; echo "debug small rc=$?"; grep -v "differing rows        0" $O/debug_small.log | grep -v "^Dimension\|^grid size\|^Number of\|^Fluid particle\|^No rigid" | cut -c1-300 | head -20
timeout 900 python -m pytest tests -q -m gpu > $O/pytest_gpu.log 2>&1
echo "exit $?" >> $O/pytest_gpu.log
grep -E "passed|failed|FAILED|exit" $O/pytest_gpu.log | head -20
run() {  # name, wmax, library
    if [ -n "$3" ]; then export SPH_B200_LIBRARY="$PWD/sph_project_b200/csrc/variants/libsph_b200_$3.so"; else unset SPH_B200_LIBRARY; fi
    SPH_B200_WMAX=$2 timeout 600 python tools/scene_step.py --scene data/scenes/dam_break_1m_dfsph.json --settle 1000 --steps 20 > $O/step_$1.json 2> $O/step_$1.err
    python - $O/step_$1.json $1 <<'PY'
import json, sys
try:
    d = json.load(open(sys.argv[1]))
    print(sys.argv[2], "ms/step", round(d["ms_per_step"], 3), "iters", d["stats"]["total_dfsph_iterations"], "kernel ms", round(d["kernel_ms_per_step"], 3))
    for k in d["kernels"][:6]: print(f"  {k['name']:48s} {k['launches_per_step']:6.2f} x {k['ms_per_launch']*1e3:8.1f} us  {k['share']:.3f}")
except Exception as e:
    print(sys.argv[2], "failed", e)
PY
    unset SPH_B200_LIBRARY
}
run default 2112 ""
run b444 2496 b444
run dc_ilp2 2112 dc_ilp2
NCU="ncu --profile-from-start off --set full --import-source on --clock-control none"
timeout 900 $NCU -k regex:'kb_dfsph_correct|kb_dfsph_density_change' -c 2 -f -o $O/ncu_brick_iter \
    python tools/scene_step.py --scene data/scenes/dam_break_1m_dfsph.json --settle 1000 --steps 1 --cuda-profiler --no-profile-pass > $O/ncu_brick_iter.log 2>&1
timeout 900 $NCU -k regex:'kb_build' -c 1 -f -o $O/ncu_brick_build \
    python tools/scene_step.py --scene data/scenes/dam_break_1m_dfsph.json --settle 1000 --steps 1 --cuda-profiler --no-profile-pass > $O/ncu_brick_build.log 2>&1
for r in ncu_brick_iter ncu_brick_build; do
    [ -f $O/$r.ncu-rep ] && ncu -i $O/$r.ncu-rep --page raw --csv > $O/${r}_raw.csv 2>/dev/null
done
python profiles/tools/ncu_summary.py $O/ncu_brick_iter_raw.csv $O/ncu_brick_build_raw.csv
