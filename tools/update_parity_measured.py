"""tests/golden/parity_measured.json <- the outlier fractions of a full `pytest -m gpu` run on B200
(gpurun_out/parity_metrics.json): tests/parity.py then asserts min(nominal allowance, 3 x measured, >= 10 elements).

  python tools/update_parity_measured.py [gpurun_out/parity_metrics.json]
Existing entries are kept at the LARGER of the old and the new measurement (two boxes may flip a different handful of
thresholds)."""
import json
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
src = Path(sys.argv[1]) if len(sys.argv) > 1 else ROOT / "gpurun_out" / "parity_metrics.json"
dst = ROOT / "tests" / "golden" / "parity_measured.json"
new = json.loads(src.read_text())
old = json.loads(dst.read_text()) if dst.exists() else {}
out = dict(old)
for name, m in new.items():
    if isinstance(m, dict) and "frac_over_tol" in m:
        out[name] = max(float(m["frac_over_tol"]), float(old.get(name, 0.0)))
dst.write_text(json.dumps(out, indent=0, sort_keys=True) + "\n")
print(f"{len(out)} comparisons in {dst} ({len(out) - len(old)} new)")
