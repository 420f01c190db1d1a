"""Drop-in driver with the reference's CLI and on-disk outputs (reference: run_simulation.py).

    python run_simulation.py --scene_file data/scenes/dam_break_8k_wcsph.json

Same loop as upstream — build container + solver from `simulationMethod`, `prepare()`, then
`step()` / `copy_to_vis_buffer()` per round and every `output_interval` rounds export
`{scene}_output/{cnt:06}/particle_object_{id}.ply` (ASCII PLY, as ti.tools.PLYWriter.export_ascii
writes it) — without Taichi: the GGUI window is a headless stub (no Vulkan on a compute node), and
the frame export `raw_view.png` is skipped.  `--max_rounds` bounds the run for smoke tests.
"""
import argparse
import os

import numpy as np

from SPH.containers import DFSPHContainer, PCISPHContainer, WCSPHContainer
from SPH.fluid_solvers import DFSPHSolver, PCISPHSolver, WCSPHSolver
from SPH.utils import SimConfig


class HeadlessWindow:
    """Stands in for ti.ui.Window('SPH', ..., show_window=False) (run_simulation.py:70)."""
    running = True

    def save_image(self, path):
        pass


def write_ply_ascii(path, positions):
    """ASCII PLY with x, y, z vertex properties (ti.tools.PLYWriter.add_vertex_pos + export_ascii)."""
    positions = np.asarray(positions, dtype=np.float32)
    with open(path, "w") as f:
        f.write("ply\nformat ascii 1.0\ncomment created by sph_project_b200\n")
        f.write(f"element vertex {positions.shape[0]}\nproperty float x\nproperty float y\nproperty float z\nend_header\n")
        np.savetxt(f, positions, fmt="%.9g")


SOLVERS = {"dfsph": (DFSPHContainer, DFSPHSolver), "wcsph": (WCSPHContainer, WCSPHSolver),
           "pcisph": (PCISPHContainer, PCISPHSolver)}

if __name__ == "__main__":
    parser = argparse.ArgumentParser()
    parser.add_argument("--scene_file", default="", help="scene file")
    parser.add_argument("--max_rounds", type=int, default=None, help="stop after this many steps (not upstream)")
    args = parser.parse_args()
    scene_path = args.scene_file
    config = SimConfig(scene_file_path=scene_path)
    scene_name = scene_path.split("/")[-1].split(".")[0]

    output_frames = config.get_cfg("exportFrame")
    fps = config.get_cfg("fps")
    if fps is None:
        fps = 60
    frame_time = 1.0 / fps
    output_interval = int(frame_time / config.get_cfg("timeStepSize"))
    total_time = config.get_cfg("totalTime")
    if total_time is None:
        total_time = 10.0
    total_rounds = int(total_time / config.get_cfg("timeStepSize"))
    if config.get_cfg("outputInterval"):
        output_interval = config.get_cfg("outputInterval")
    if args.max_rounds is not None:
        total_rounds = min(total_rounds, args.max_rounds)
    output_ply = config.get_cfg("exportPly")
    output_obj = config.get_cfg("exportObj")

    os.makedirs(f"{scene_name}_output", exist_ok=True)

    simulation_method = config.get_cfg("simulationMethod")
    if simulation_method not in SOLVERS:
        # iisph / pbf are broken upstream (README.md:11,215-216) and not part of the hot path
        raise NotImplementedError(f"Simulation method {simulation_method} not implemented")
    Container, Solver = SOLVERS[simulation_method]
    container = Container(config, GGUI=True)
    solver = Solver(container)
    print(f"Simulation method: {simulation_method}")

    solver.prepare()
    window = HeadlessWindow()

    invisible_objects = config.get_cfg("invisibleObjects") or []
    dim = len(config.get_cfg("domainEnd"))

    cnt = 0
    while window.running:
        solver.step()
        if cnt % output_interval == 0:
            if output_frames:
                container.copy_to_vis_buffer(invisible_objects=invisible_objects, dim=dim)
                os.makedirs(f"{scene_name}_output/{cnt:06}", exist_ok=True)
                window.save_image(f"{scene_name}_output/{cnt:06}/raw_view.png")
            if output_ply:
                os.makedirs(f"{scene_name}_output/{cnt:06}", exist_ok=True)
                for f_body_id in container.object_id_fluid_body:
                    obj_data = container.dump(obj_id=f_body_id)
                    write_ply_ascii(f"{scene_name}_output/{cnt:06}/particle_object_{f_body_id}.ply", obj_data["position"])
            if output_obj:
                os.makedirs(f"{scene_name}_output/{cnt:06}", exist_ok=True)
                for r_body_id in container.object_id_rigid_body:
                    mesh = container.object_collection[r_body_id].get("mesh")
                    if mesh is not None:
                        with open(f"{scene_name}_output/{cnt:06}/mesh_object_{r_body_id}.obj", "w") as f:
                            f.write(mesh.export(file_type="obj"))
        cnt += 1
        if cnt >= total_rounds:
            break

    print("Simulation Finished")
