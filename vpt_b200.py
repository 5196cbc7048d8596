"""Import shim: the package directory is named `vulkan-path-tracer_b200` (not a valid Python identifier),
so `import vpt_b200` loads it under that alias."""
import importlib.util as _u
import os as _os
import sys as _sys

_dir = _os.path.join(_os.path.dirname(_os.path.abspath(__file__)), "vulkan-path-tracer_b200")
_spec = _u.spec_from_file_location("vpt_b200_pkg", _os.path.join(_dir, "__init__.py"), submodule_search_locations=[_dir])
_mod = _u.module_from_spec(_spec)
_sys.modules["vpt_b200_pkg"] = _mod
_spec.loader.exec_module(_mod)
binding = _mod.binding
globals().update({k: getattr(binding, k) for k in dir(binding) if not k.startswith("_")})
