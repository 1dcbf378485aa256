"""`kungfu` — import-name compatibility with the reference (`import kungfu.torch as kf`, `kungfu.python`, `kungfu.cmd`).
Everything lives in `kungfu_b200`; this package only aliases the modules."""
import importlib
import sys

import kungfu_b200 as _impl

__version__ = _impl.__version__

for _name in ("python", "torch", "cmd", "ops", "elastic", "initializer", "checkpoint", "utils", "models", "parallel", "helpers", "platforms",
              "benchmarks", "variables", "examples"):
    _mod = importlib.import_module(f"kungfu_b200.{_name}")
    sys.modules[f"kungfu.{_name}"] = _mod
    globals()[_name] = _mod
for _sub in ("torch.ops", "torch.optimizers"):
    sys.modules[f"kungfu.{_sub}"] = importlib.import_module(f"kungfu_b200.{_sub}")
sys.modules["kungfu._utils"] = sys.modules["kungfu.utils"]  # reference: kungfu/_utils.py (measure, show_duration, _log_event)
_utils = sys.modules["kungfu.utils"]
sys.modules["kungfu.info"] = importlib.import_module("kungfu_b200.info")
info = sys.modules["kungfu.info"]  # reference: `python -m kungfu.info`
