"""Import shim: the per-period distance table of scripts/golden_drift.py, for tests."""
import importlib.util
import os

_p = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "scripts", "golden_drift.py")
_spec = importlib.util.spec_from_file_location("golden_drift", _p)
_m = importlib.util.module_from_spec(_spec)
_spec.loader.exec_module(_m)
per_period_table = _m.per_period_table
