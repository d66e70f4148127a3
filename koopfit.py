"""Import shim: `import koopfit` loads the package in `koopman-realizations_b200/`
(a directory name Python cannot import directly because of the hyphen)."""
import importlib.util
import os
import sys

_dir = os.path.join(os.path.dirname(os.path.abspath(__file__)), "koopman-realizations_b200")
_spec = importlib.util.spec_from_file_location(
    "koopfit", os.path.join(_dir, "__init__.py"), submodule_search_locations=[_dir])
_mod = importlib.util.module_from_spec(_spec)
sys.modules["koopfit"] = _mod
_spec.loader.exec_module(_mod)
