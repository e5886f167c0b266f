"""Import alias: `import b200_whisper` loads the package that lives in the (hyphenated, hence not directly
importable) directory eddie-wang-hackathon2023_b200/."""
import importlib.util
import os
import sys

_pkg_dir = os.path.join(os.path.dirname(os.path.abspath(__file__)), "eddie-wang-hackathon2023_b200")
_spec = importlib.util.spec_from_file_location(
    "b200_whisper", os.path.join(_pkg_dir, "__init__.py"), submodule_search_locations=[_pkg_dir])
_mod = importlib.util.module_from_spec(_spec)
sys.modules["b200_whisper"] = _mod
_spec.loader.exec_module(_mod)
