"""Import shim: the package directory is named `heterogeneous-spmv_b200` (with a hyphen, as the
task layout prescribes), which Python cannot import by name.  `import heterogeneous_spmv_b200`
loads that directory as a regular package (sub-modules such as `.generators` included)."""
import importlib.util
import os
import sys

_dir = os.path.join(os.path.dirname(os.path.abspath(__file__)), "heterogeneous-spmv_b200")
_spec = importlib.util.spec_from_file_location(
    __name__, os.path.join(_dir, "__init__.py"), submodule_search_locations=[_dir])
_mod = importlib.util.module_from_spec(_spec)
sys.modules[__name__] = _mod
_spec.loader.exec_module(_mod)
