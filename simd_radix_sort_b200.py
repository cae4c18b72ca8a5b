"""Import shim: the package directory is named `simd-radix-sort_b200` (not an identifier), so
`import simd_radix_sort_b200` loads it from there."""
import importlib.util
import sys
from pathlib import Path

_pkg = Path(__file__).resolve().parent / "simd-radix-sort_b200"
_spec = importlib.util.spec_from_file_location("simd_radix_sort_b200", _pkg / "__init__.py",
                                               submodule_search_locations=[str(_pkg)])
_mod = importlib.util.module_from_spec(_spec)
sys.modules["simd_radix_sort_b200"] = _mod
_spec.loader.exec_module(_mod)
