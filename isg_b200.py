"""Import shim: the package directory is `instance-segmentation_b200/` (not a valid Python identifier), so
`import isg_b200` loads that directory as the package `isg_b200` (submodules resolve normally, e.g.
`import isg_b200.utils.decode`)."""
import importlib.util as _ilu
import os as _os
import sys as _sys

_dir = _os.path.join(_os.path.dirname(_os.path.abspath(__file__)), "instance-segmentation_b200")
_spec = _ilu.spec_from_file_location("isg_b200", _os.path.join(_dir, "__init__.py"), submodule_search_locations=[_dir])
_mod = _ilu.module_from_spec(_spec)
_sys.modules["isg_b200"] = _mod
_spec.loader.exec_module(_mod)
