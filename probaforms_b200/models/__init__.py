"""``from probaforms_b200.models import RealNVP`` mirrors ``from probaforms.models import RealNVP``
(reference probaforms/models/__init__.py:1).  Only the RealNVP path is in scope."""
from .interfaces import GenModel
from .nflow import InvertibleLayer, NormalizingFlow
from .realnvp import RealNVP, RealNVPLayer, gen_network

__all__ = ["RealNVP", "RealNVPLayer", "NormalizingFlow", "InvertibleLayer", "GenModel", "gen_network"]
