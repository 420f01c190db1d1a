from .._native import F
from ..fields import ParticleField
from .base_container import BaseContainer


class DFSPHContainer(BaseContainer):
    """DFSPH scratch fields (reference: containers/dfsph_container.py:13-17)."""
    _method = "dfsph"

    def __init__(self, config, GGUI=False, **kw):
        super().__init__(config, GGUI, **kw)
        eng, cap = self._engine, self.particle_max_num
        self.particle_dfsph_alphas = ParticleField(eng, F.DFSPH_ALPHA, cap)
        self.particle_dfsph_kappa = ParticleField(eng, F.DFSPH_KAPPA, cap)
        self.particle_dfsph_kappa_v = ParticleField(eng, F.DFSPH_KAPPA_V, cap)
        self.particle_densities_star = ParticleField(eng, F.DENSITY_STAR, cap)
        self.particle_densities_derivatives = ParticleField(eng, F.DENSITY_DERIVATIVE, cap)
