"""Scene configuration reader (reference: SPH/utils/config_builder.py:5-44).

Same constructor and getters as the reference's SimConfig: a scene is one JSON file with a
"Configuration" dict plus optional "FluidBlocks" / "FluidBodies" / "RigidBodies" /
"RigidBlocks" lists; missing keys read as None, missing lists as [].
"""
import json

_BODY_LISTS = ("RigidBodies", "RigidBlocks", "FluidBodies", "FluidBlocks")


class SimConfig:
    def __init__(self, scene_file_path=None, config=None, verbose=True) -> None:
        if config is None:
            with open(scene_file_path, "r") as fh:
                config = json.load(fh)
        self.config = config
        if verbose:
            print(self.config)

    def get_cfg(self, name, enforce_exist=False):
        section = self.config["Configuration"]
        if name in section:
            return section[name]
        assert not enforce_exist, f"scene file has no Configuration.{name}"
        return None

    def _list(self, key):
        return self.config.get(key, [])

    def get_rigid_bodies(self):
        return self._list("RigidBodies")

    def get_rigid_blocks(self):
        return self._list("RigidBlocks")

    def get_fluid_bodies(self):
        return self._list("FluidBodies")

    def get_fluid_blocks(self):
        return self._list("FluidBlocks")
