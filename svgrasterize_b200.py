"""Import shim: ``import svgrasterize_b200`` loads the package that lives in the
directory ``svgrasterize.py_b200/`` (a dotted directory name is not importable
with a plain ``import`` statement, so this module replaces itself in
``sys.modules`` with the package object)."""
import importlib.util
import os
import sys

_dir = os.path.join(os.path.dirname(os.path.abspath(__file__)), "svgrasterize.py_b200")
_spec = importlib.util.spec_from_file_location(
    __name__, os.path.join(_dir, "__init__.py"), submodule_search_locations=[_dir]
)
_mod = importlib.util.module_from_spec(_spec)
sys.modules[__name__] = _mod
_spec.loader.exec_module(_mod)
