"""Import shim: the package directory `node-fhe-accelerate_b200` is not a valid Python
identifier, so `import fheb200` loads it under the module name `node_fhe_accelerate_b200`."""
import importlib.util
import os
import sys

_NAME = "node_fhe_accelerate_b200"
_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "node-fhe-accelerate_b200")

if _NAME not in sys.modules:
    _spec = importlib.util.spec_from_file_location(_NAME, os.path.join(_DIR, "__init__.py"),
                                                   submodule_search_locations=[_DIR])
    _mod = importlib.util.module_from_spec(_spec)
    sys.modules[_NAME] = _mod
    _spec.loader.exec_module(_mod)

_pkg = sys.modules[_NAME]
globals().update({k: v for k, v in vars(_pkg).items() if not k.startswith("__")})
