from .._native import F, S
from ..fields import ParticleField, ScalarField
from .base_container import BaseContainer


class PCISPHContainer(BaseContainer):
    """PCISPH scratch fields (reference: containers/pcisph_container.py:14-19)."""
    _method = "pcisph"

    def __init__(self, config, GGUI=False, **kw):
        super().__init__(config, GGUI, **kw)
        eng, cap = self._engine, self.particle_max_num
        self.density_error = ScalarField(eng, S.DENSITY_ERROR)
        self.pcisph_k = ScalarField(eng, S.PCISPH_K)
        self.particle_pressure_accelerations = ParticleField(eng, F.PRESSURE_ACCELERATION, cap)
        self.particle_predicted_velocities = ParticleField(eng, F.PREDICTED_VELOCITY, cap)
        self.particle_predicted_positions = ParticleField(eng, F.PREDICTED_POSITION, cap)
        self.particle_densities_star = ParticleField(eng, F.DENSITY_STAR, cap)
