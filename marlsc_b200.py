"""Import alias: the package lives in ``marl-sc_b200/`` (not an importable name), so it is loaded
here once under the module name ``marlsc_b200``; ``import marlsc_b200.envs`` etc. then resolve
through its ``__path__``."""
import importlib.util as _u
import os as _os
import sys as _sys

_dir = _os.path.join(_os.path.dirname(_os.path.abspath(__file__)), "marl-sc_b200")
_spec = _u.spec_from_file_location("marlsc_b200", _os.path.join(_dir, "__init__.py"),
                                   submodule_search_locations=[_dir])
_mod = _u.module_from_spec(_spec)
_sys.modules["marlsc_b200"] = _mod
_spec.loader.exec_module(_mod)
