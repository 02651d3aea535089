"""Loader: makes the package directory `highvoronoi.jl_b200/` importable as `hvb200`."""
import importlib.util
import os
import sys

_d = os.path.join(os.path.dirname(os.path.abspath(__file__)), "highvoronoi.jl_b200")
_spec = importlib.util.spec_from_file_location("hvb200", os.path.join(_d, "__init__.py"), submodule_search_locations=[_d])
_mod = importlib.util.module_from_spec(_spec)
sys.modules["hvb200"] = _mod
_spec.loader.exec_module(_mod)
