"""Import shim: `import viet_asr_b200` loads the package that lives in ./viet-asr_b200/
(the directory name the project layout prescribes is not a valid Python identifier)."""
import importlib.util
import os
import sys

_real = os.path.join(os.path.dirname(os.path.abspath(__file__)), "viet-asr_b200")
_spec = importlib.util.spec_from_file_location(
    "viet_asr_b200", os.path.join(_real, "__init__.py"), submodule_search_locations=[_real])
_mod = importlib.util.module_from_spec(_spec)
sys.modules["viet_asr_b200"] = _mod
_spec.loader.exec_module(_mod)
