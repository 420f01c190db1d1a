#!/bin/bash
# Round-2 GPU call 1 (1 x B200): every GPU test incl. the former gpu_next ones, ncu --set full of the kernels the
# metric names (density / list build, pressure force, sort gather) on C2 (WCSPH) and C2' (DFSPH, pressurised),
# per-scene step timings of C2 / C3 / C4, and the compile-time variant sweep of the round-1 kernels.
O=gpurun_out/r02c1
mkdir -p $O
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $O/nvidia_smi.csv 2>&1
SPH_RUN_GPU_NEXT=1 timeout 900 python -m pytest tests -q -m gpu > $O/pytest_gpu_all.log 2>&1
echo "exit $?" >> $O/pytest_gpu_all.log
tail -5 $O/pytest_gpu_all.log
for sc in dam_break_1m_wcsph:300 bath_500k_dfsph:300 buckling_pcisph_implicit:50 dam_break_1m_dfsph:1000; do
    name=${sc%%:*}; settle=${sc##*:}
    timeout 600 python tools/scene_step.py --scene data/scenes/$name.json --settle $settle --steps 20 > $O/step_$name.json 2> $O/step_$name.err
    echo "$name rc=$?"; head -c 400 $O/step_$name.json; echo
done
NCU="ncu --profile-from-start off --set full --import-source on --clock-control none"
timeout 900 $NCU -k regex:'k_density|k_pressure_accel|k_gather|k_viscosity|k_surface_tension' -c 5 -f -o $O/ncu_c2_wcsph \
    python tools/scene_step.py --scene data/scenes/dam_break_1m_wcsph.json --settle 300 --steps 1 --cuda-profiler --no-profile-pass > $O/ncu_c2_wcsph.log 2>&1
timeout 900 $NCU -k regex:'k_density|k_dfsph_alpha|k_gather' -c 3 -f -o $O/ncu_c2p_dfsph \
    python tools/scene_step.py --scene data/scenes/dam_break_1m_dfsph.json --settle 1000 --steps 1 --cuda-profiler --no-profile-pass > $O/ncu_c2p_dfsph.log 2>&1
for r in ncu_c2_wcsph ncu_c2p_dfsph; do
    [ -f $O/$r.ncu-rep ] && ncu -i $O/$r.ncu-rep --page raw --csv > $O/${r}_raw.csv 2>/dev/null
done
ls -la $O
LIBDIR=sph_project_b200/csrc
for name in default minb10 unroll8 block256 all_hints devconv; do
    if [ "$name" = default ]; then unset SPH_B200_LIBRARY; else export SPH_B200_LIBRARY="$PWD/$LIBDIR/variants/libsph_b200_$name.so"; fi
    [ "$name" = default ] || [ -f "$SPH_B200_LIBRARY" ] || { echo "$name: not built"; continue; }
    timeout 300 python -m pytest -q -m gpu -x "tests/test_gpu_parity.py::test_trajectory_parity" "tests/test_gpu_parity.py::test_pressurised_dfsph_iterations_match" > "$O/variant_$name.parity.log" 2>&1
    echo "parity exit $?" >> "$O/variant_$name.parity.log"
    timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > "$O/variant_$name.json" 2> "$O/variant_$name.err"
    python - "$O/variant_$name.json" "$name" "$(tail -1 $O/variant_$name.parity.log)" <<'PY'
import json, sys
try:
    d = json.loads(open(sys.argv[1]).read())
    ks = ", ".join(f"{k['name'].split('<')[0][2:]} {k['ms_per_launch']:.3f}" for k in d["roofline"]["kernels"][:4])
    print(f"{sys.argv[2]:12s} {sys.argv[3]:14s} {d['value'] / 1e6:8.1f} M  {d['ms_per_step']:7.2f} ms  it {d['config']['mean_iterations']}  {ks}")
except Exception as e:
    print(sys.argv[2], "no bench line", e)
PY
done
unset SPH_B200_LIBRARY
