"""Import shim: the package directory is named ``polgen-rvc_b200`` (hyphen, per
the repo layout contract), which Python cannot import by name.  Importing
``polgen_rvc_b200`` loads that directory as a regular package under this name.
"""
import importlib.util as _ilu
import os as _os
import sys as _sys

_pkg_dir = _os.path.join(_os.path.dirname(_os.path.abspath(__file__)), "polgen-rvc_b200")
_spec = _ilu.spec_from_file_location(
    "polgen_rvc_b200", _os.path.join(_pkg_dir, "__init__.py"),
    submodule_search_locations=[_pkg_dir])
_mod = _ilu.module_from_spec(_spec)
_sys.modules["polgen_rvc_b200"] = _mod
_spec.loader.exec_module(_mod)
