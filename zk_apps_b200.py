"""Import shim: the package directory is `zk-apps_b200/` (the name the task fixes), which is not a
valid Python identifier; `import zk_apps_b200` loads it from there."""
import importlib.util
import os
import sys

_dir = os.path.join(os.path.dirname(os.path.abspath(__file__)), "zk-apps_b200")
_spec = importlib.util.spec_from_file_location("zk_apps_b200", os.path.join(_dir, "__init__.py"),
                                               submodule_search_locations=[_dir])
_mod = importlib.util.module_from_spec(_spec)
sys.modules["zk_apps_b200"] = _mod
_spec.loader.exec_module(_mod)
