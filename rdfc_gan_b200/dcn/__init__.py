"""Mirror of the reference's vendored ``deformconv`` package (modules + functions + the ``DCN`` extension)."""
from . import DCN
from .functions import DeformConvFunction, ModulatedDeformConvFunction
from .modules import (DeformConv, DeformConvPack, ModulatedDeformConv, ModulatedDeformConvPack, _DeformConv,
                      _ModulatedDeformConv)

__all__ = ["DCN", "DeformConvFunction", "ModulatedDeformConvFunction", "DeformConv", "DeformConvPack",
           "ModulatedDeformConv", "ModulatedDeformConvPack", "_DeformConv", "_ModulatedDeformConv"]
