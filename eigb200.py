"""Import shim: `import eigb200` loads the package that lives in
`task-level-insights-from-eigenvalues-across-sequence-models_b200/` (a directory name Python cannot import directly)."""
import importlib.util
import os
import sys

_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)),
                    "task-level-insights-from-eigenvalues-across-sequence-models_b200")
_spec = importlib.util.spec_from_file_location("eigb200", os.path.join(_DIR, "__init__.py"),
                                               submodule_search_locations=[_DIR])
_mod = importlib.util.module_from_spec(_spec)
sys.modules["eigb200"] = _mod
_spec.loader.exec_module(_mod)
