"""Import shim: loads the package directory `factor-fields_b200/` under the importable name `ffb200`."""
import importlib.util
import os
import sys

_dir = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'factor-fields_b200')
_spec = importlib.util.spec_from_file_location('ffb200', os.path.join(_dir, '__init__.py'), submodule_search_locations=[_dir])
_mod = importlib.util.module_from_spec(_spec)
sys.modules['ffb200'] = _mod
_spec.loader.exec_module(_mod)
