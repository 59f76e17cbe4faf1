"""Import shim: the package directory is `raytracer-2022_b200` (not a valid Python identifier).

    import rtb200 as rt
    scene = rt.Scene(seed=2022).build_config("final_scene")
"""
import importlib.util
import os
import sys

_NAME = "raytracer_2022_b200"
if _NAME not in sys.modules:
    _dir = os.path.join(os.path.dirname(os.path.abspath(__file__)), "raytracer-2022_b200")
    _spec = importlib.util.spec_from_file_location(_NAME, os.path.join(_dir, "__init__.py"), submodule_search_locations=[_dir])
    _mod = importlib.util.module_from_spec(_spec)
    sys.modules[_NAME] = _mod
    _spec.loader.exec_module(_mod)
_pkg = sys.modules[_NAME]
globals().update({k: v for k, v in vars(_pkg).items() if not k.startswith("__")})
D = _pkg.D
