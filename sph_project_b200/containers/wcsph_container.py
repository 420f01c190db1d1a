from .base_container import BaseContainer


class WCSPHContainer(BaseContainer):
    """WCSPH needs no fields beyond the base set (reference: containers/wcsph_container.py:10-12)."""
    _method = "wcsph"
