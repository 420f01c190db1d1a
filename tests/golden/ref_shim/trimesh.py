"""Import stub: the reference imports trimesh at module level; mesh bodies are not used by the fixture scenes."""


def load(*a, **k):
    raise NotImplementedError("trimesh is not available offline; fixture scenes use FluidBlocks + the domain box only")
