"""zodipy_b200: B200-native line-of-sight brightness integration behind ZodiPy's ``Model`` API."""
from .model import Model
from .units import Quantity
from .zodiacal_light_model import model_registry

__all__ = ("Model", "Quantity", "model_registry")
__version__ = "0.1.0"
