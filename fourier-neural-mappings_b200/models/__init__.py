"""Model zoo with the reference's names (reference: models/__init__.py:1-11)."""
from .cross_dim import FNO1d2, FNO2d1
from .func_to_func import FNO1d, FNO2d
from .func_to_vec import FNF1d, FNF2d
from .vec_to_func import FND1d, FND2d
from .vec_to_vec import FNN1d, FNN2d

__all__ = ["FNO1d", "FNO2d", "FNF1d", "FNF2d", "FND1d", "FND2d", "FNN1d", "FNN2d", "FNO1d2", "FNO2d1"]
