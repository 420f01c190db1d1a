from .config_builder import SimConfig

__all__ = ["SimConfig"]
