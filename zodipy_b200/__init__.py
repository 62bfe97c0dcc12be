"""zodipy_b200: B200-native line-of-sight brightness integration behind ZodiPy's ``Model`` API."""
from .model import Model, MultiBandModel
from .number_density import grid_number_density, grid_number_density_xyz
from .units import Quantity
from .zodiacal_light_model import model_registry

__all__ = ("Model", "MultiBandModel", "Quantity", "grid_number_density", "grid_number_density_xyz", "model_registry")
__version__ = "0.1.0"
