"""hvb200 -- B200-native raycast vertex-search backend for HighVoronoi.jl (host-side mirror of the reference API).

The directory is called `highvoronoi.jl_b200`; because of the dot it is imported through the loader module
`hvb200.py` at the repository root (`import hvb200`)."""
from . import _abi  # noqa: F401
from .api import (refine, ConvexHull, B200Thread, Boundary, HVBError, Raycast, RaycastParameter, RCCombined, RCNonGeneral,  # noqa: F401
                  RCNonGeneralFast, RCNonGeneralHP, RCOriginal, RCStandard, RCOriginalSafety, RCNonGeneralSkip, RCOriginalHP, RCNonGeneralCutoff, SingleThread, VoronoiData,
                  VoronoiGeometry, VoronoiMesh, VoronoiNodes, cuboid, voronoi)

__version__ = "0.2.0"
