"""Import shim: the package directory is named ``reachability.jl_b200`` (not an importable identifier), so
``import reachability_jl_b200`` loads it from there under this name."""
import importlib.util as _ilu
import pathlib as _pl
import sys as _sys

_dir = _pl.Path(__file__).resolve().parent / "reachability.jl_b200"
_spec = _ilu.spec_from_file_location("reachability_jl_b200", _dir / "__init__.py",
                                     submodule_search_locations=[str(_dir)])
_mod = _ilu.module_from_spec(_spec)
_sys.modules["reachability_jl_b200"] = _mod
_spec.loader.exec_module(_mod)
