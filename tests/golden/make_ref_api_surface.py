"""Public attribute surface of the reference's container / solver objects (instantiated on the Taichi emulation,
see make_ref_golden.py) -> tests/golden/ref_api_surface.json.  tests/test_api_surface.py checks that this
repository's drop-in classes offer the same names.

    python tests/golden/make_ref_api_surface.py
"""
import contextlib
import io
import json
import os
import sys
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import make_ref_golden as G  # noqa: E402


def kind(obj, name):
    try:
        v = getattr(obj, name)
    except Exception:      # noqa: BLE001
        return "unreadable"
    if callable(v):
        # ti.func = device-side only upstream (calling it from Python scope raises in Taichi); ti.kernel and plain
        # methods are the host-callable surface
        return {"kernel": "ti.kernel", "func": "ti.func"}.get(getattr(v, "__ti_kind__", None), "method")
    if hasattr(v, "to_numpy"):
        return "field"
    return "value"


def surface(obj):
    names = set(vars(obj)) | {n for n in dir(type(obj)) if not n.startswith("__")}
    return {n: kind(obj, n) for n in sorted(names)}


def main():
    SimConfig, classes = G.import_reference()
    out = {}
    for method in ("wcsph", "pcisph", "dfsph"):
        sc = G.scene(method=method)
        with tempfile.NamedTemporaryFile("w", suffix=".json", delete=False) as fh:
            json.dump(sc, fh)
        with contextlib.redirect_stdout(io.StringIO()):
            cfg = SimConfig(scene_file_path=fh.name)
            C, S = classes[method]
            container = C(cfg, GGUI=True)
            solver = S(container)
        os.unlink(fh.name)
        out[method] = {"container": surface(container), "solver": surface(solver), "rigid_solver": surface(solver.rigid_solver),
                       "config": surface(cfg)}
    with open(os.path.join(HERE, "ref_api_surface.json"), "w") as fh:
        json.dump(out, fh, indent=1, sort_keys=True)
    print({m: {k: len(v) for k, v in d.items()} for m, d in out.items()})


if __name__ == "__main__":
    main()
