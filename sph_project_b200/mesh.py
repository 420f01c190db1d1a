"""Mesh bodies -> particles (reference base_container.py:611-717).

The reference voxelises OBJ meshes with trimesh, which this image does not have.  Mesh bodies are a
"next" row of the hot-path scope table; until the voxeliser lands, scenes with FluidBodies /
RigidBodies must be run with those lists removed (as BASELINE.md's C2'/C3 configurations do).
"""


def voxelize_rigid_body(rigid_body, pitch):
    raise NotImplementedError(
        f"RigidBodies need the mesh voxeliser (not built yet): {rigid_body.get('geometryFile')}")


def voxelize_fluid_body(fluid_body, pitch, dim):
    raise NotImplementedError(
        f"FluidBodies need the mesh voxeliser (not built yet): {fluid_body.get('geometryFile')}")
