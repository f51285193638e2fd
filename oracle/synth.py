"""The synthetic inputs of SURVEY.md section 8(d) live in the product package (neural-color-transfer_b200/synth.py: pure
numpy, used by bench.py's product arm, which may not import oracle/); this shim re-exports them for the tests and the
oracle, which may import anything."""
import importlib.util
import os
import sys

_path = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "neural-color-transfer_b200", "synth.py")
_spec = importlib.util.spec_from_file_location("nct_b200_synth", _path)
_mod = sys.modules.get("nct_b200_synth")
if _mod is None:
    _mod = importlib.util.module_from_spec(_spec)
    sys.modules["nct_b200_synth"] = _mod
    _spec.loader.exec_module(_mod)
globals().update({k: v for k, v in vars(_mod).items() if not k.startswith("__")})
