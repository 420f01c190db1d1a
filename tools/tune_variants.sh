#!/bin/bash
# Tuning sweep over the pre-built variant libraries (make -C sph_project_b200/csrc variants), one gpurun call:
#   make -C sph_project_b200/csrc variants        # here, before the call: nvcc cross-compiles without a GPU
#   /usr/local/graft/bin/gpurun --timeout 1500 -- 'bash tools/tune_variants.sh'
# Every variant must reproduce the window walk bit for bit (same arithmetic, same summation order) before its
# bench line counts.  Results: gpurun_out/variant_<name>.json (+ .err, .parity.log), table on stdout.
mkdir -p gpurun_out
LIBDIR=sph_project_b200/csrc
for name in default b442 b442_256 b442_320 b342_256; do   # the VARIANTS of sph_project_b200/csrc/Makefile (window budgets: see tools/r02_call20.sh)
    if [ "$name" = default ]; then unset SPH_B200_LIBRARY; else export SPH_B200_LIBRARY="$PWD/$LIBDIR/variants/libsph_b200_$name.so"; fi
    [ "$name" = default ] || [ -f "$SPH_B200_LIBRARY" ] || { echo "$name: not built"; continue; }
    timeout 300 python -m pytest -q -m gpu -x "tests/test_gpu_fullsize.py::test_list_kernels_equal_window_walk_bitwise" \
        "tests/test_gpu_parity.py::test_trajectory_parity" > "gpurun_out/variant_$name.parity.log" 2>&1
    echo "parity exit $?" >> "gpurun_out/variant_$name.parity.log"
    timeout 300 python bench.py --steps 30 --warmup 5 --no-cpu-baseline > "gpurun_out/variant_$name.json" 2> "gpurun_out/variant_$name.err"
done
unset SPH_B200_LIBRARY
python - <<'PY'
import glob, json, os
print(f"{'variant':18s} {'parity':>7s} {'M pps':>8s} {'ms/step':>8s}  top kernels (ms per launch)")
for path in sorted(glob.glob("gpurun_out/variant_*.json")):
    name = os.path.basename(path)[8:-5]
    parity = open(path[:-5] + ".parity.log").read().strip().splitlines()[-1] if os.path.exists(path[:-5] + ".parity.log") else "?"
    try:
        d = json.loads(open(path).read())
    except ValueError:
        print(f"{name:18s} {parity:>7s}  no bench line"); continue
    ks = ", ".join(f"{k['name'].split('<')[0][2:]} {k['ms_per_launch']:.3f}" for k in d["roofline"]["kernels"][:3])
    print(f"{name:18s} {parity.split()[-1]:>7s} {d['value'] / 1e6:8.1f} {d['ms_per_step']:8.2f}  {ks}")
PY
