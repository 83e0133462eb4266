"""Import alias for the package directory `visual-slam-core_b200/`.

The directory name (fixed by the build contract) is not a Python identifier, so
this shim loads it under the module name ``visual_slam_core_b200``:

    import b200ba                       # registers the package
    from visual_slam_core_b200 import B200Optimizer
"""
import importlib.util
import os
import sys

_NAME = "visual_slam_core_b200"
_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "visual-slam-core_b200")

if _NAME not in sys.modules:
    _spec = importlib.util.spec_from_file_location(
        _NAME, os.path.join(_DIR, "__init__.py"), submodule_search_locations=[_DIR])
    _mod = importlib.util.module_from_spec(_spec)
    sys.modules[_NAME] = _mod
    _spec.loader.exec_module(_mod)

package = sys.modules[_NAME]
PACKAGE_DIR = _DIR
