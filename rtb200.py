"""Import shim: the package directory is named `raytracing-course-2024_b200` (not a valid Python identifier), so
it is loaded by path and exposed as module `rtb200`:   import rtb200 as rt"""
import importlib.util as _u
import os as _os
import sys as _sys

_dir = _os.path.join(_os.path.dirname(_os.path.abspath(__file__)), "raytracing-course-2024_b200")
_name = "raytracing_course_2024_b200"
if _name not in _sys.modules:
    _spec = _u.spec_from_file_location(_name, _os.path.join(_dir, "__init__.py"), submodule_search_locations=[_dir])
    _mod = _u.module_from_spec(_spec)
    _sys.modules[_name] = _mod
    _spec.loader.exec_module(_mod)
_pkg = _sys.modules[_name]
globals().update({k: getattr(_pkg, k) for k in dir(_pkg) if not k.startswith("__")})
