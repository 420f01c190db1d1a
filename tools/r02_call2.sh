#!/bin/bash
# Round-2 GPU call 2: first hardware run of the brick-tile sweeps (correctness first, then a first timing)
O=gpurun_out/r02c2
mkdir -p $O
timeout 900 python -m pytest tests -q -m gpu -x > $O/pytest_gpu.log 2>&1
echo "exit $?" >> $O/pytest_gpu.log
tail -30 $O/pytest_gpu.log
timeout 300 python tools/scene_step.py --scene data/scenes/dam_break_1m_dfsph.json --settle 40 --steps 20 > $O/step_dfsph_early.json 2> $O/step_dfsph_early.err
echo "early rc=$?"; head -c 1500 $O/step_dfsph_early.json; echo; tail -3 $O/step_dfsph_early.err
timeout 600 python tools/scene_step.py --scene data/scenes/dam_break_1m_dfsph.json --settle 1000 --steps 20 > $O/step_dfsph_press.json 2> $O/step_dfsph_press.err
echo "press rc=$?"; head -c 3000 $O/step_dfsph_press.json; echo; tail -3 $O/step_dfsph_press.err
timeout 300 python tools/scene_step.py --scene data/scenes/dam_break_1m_wcsph.json --settle 300 --steps 20 > $O/step_wcsph.json 2> $O/step_wcsph.err
echo "wcsph rc=$?"; head -c 2500 $O/step_wcsph.json; echo
