"""Import shim: the package directory is named `build-up-phase_b200/` (after the reference repo), which
is not a valid Python identifier. `import build_up_phase_b200` loads that directory as a package."""
import importlib.util as _ilu
import os as _os
import sys as _sys

_d = _os.path.join(_os.path.dirname(_os.path.abspath(__file__)), "build-up-phase_b200")
_spec = _ilu.spec_from_file_location(__name__, _os.path.join(_d, "__init__.py"), submodule_search_locations=[_d])
_mod = _ilu.module_from_spec(_spec)
_sys.modules[__name__] = _mod
_spec.loader.exec_module(_mod)
