#!/bin/bash
# Round-2 GPU call: rolling-prefetch and 4x4x2-brick variants against the shipped build (pressurised C2')
O=gpurun_out/r02c20
mkdir -p $O
run() {  # name, wmax, library
    if [ -n "$3" ]; then export SPH_B200_LIBRARY="$PWD/sph_project_b200/csrc/variants/libsph_b200_$3.so"; else unset SPH_B200_LIBRARY; fi
    SPH_B200_WMAX=$2 timeout 300 python -m pytest -q -m gpu -x tests/test_gpu_fullsize.py::test_list_kernels_equal_window_walk_bitwise tests/test_gpu_parity.py::test_trajectory_parity > $O/parity_$1.log 2>&1; echo "$1 parity exit $?"
    SPH_B200_WMAX=$2 timeout 600 python tools/scene_step.py --scene data/scenes/dam_break_1m_dfsph.json --settle 1000 --steps 20 > $O/step_$1.json 2> $O/step_$1.err
    python - $O/step_$1.json $1 <<'PY'
import json, sys
try:
    d = json.load(open(sys.argv[1]))
    print(sys.argv[2], "ms/step", round(d["ms_per_step"], 3), "iters", d["stats"]["total_dfsph_iterations"], "kernel ms", round(d["kernel_ms_per_step"], 3))
    for k in d["kernels"][:7]: print(f"  {k['name']:48s} {k['launches_per_step']:6.2f} x {k['ms_per_launch']*1e3:8.1f} us  {k['share']:.3f}")
except Exception as e:
    print(sys.argv[2], "failed", e)
PY
    unset SPH_B200_LIBRARY
}
run b442 1664 b442
run b442_256 1664 b442_256
run b442_320 1664 b442_320
run b342_256 1280 b342_256
