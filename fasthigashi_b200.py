"""Import alias: the package lives in the directory ``fast-higashi_b200/`` (a name Python
cannot import directly), so this module loads that directory as the package
``fasthigashi_b200``.  ``import fasthigashi_b200`` therefore gives the same object a
user would get from a pip-installed copy.
"""
import os as _os
import sys as _sys
import importlib.util as _ilu

_here = _os.path.dirname(_os.path.abspath(__file__))
_pkg_dir = _os.path.join(_here, "fast-higashi_b200")
_spec = _ilu.spec_from_file_location(
	"fasthigashi_b200", _os.path.join(_pkg_dir, "__init__.py"),
	submodule_search_locations=[_pkg_dir])
_mod = _ilu.module_from_spec(_spec)
_sys.modules["fasthigashi_b200"] = _mod
_spec.loader.exec_module(_mod)
