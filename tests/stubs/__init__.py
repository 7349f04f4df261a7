"""TEST INFRASTRUCTURE: just enough of nerfstudio 1.1.3 / torchmetrics to import and run the reference's OWN
`dn_splatter/dn_model.py`, unmodified, on top of this repository's `gsplat` shim.

Neither package is installed in this image.  The stubs restate, from SURVEY.md Appendix A.7, only the
pieces `DNSplatterModel` inherits or calls: `SplatfactoModel` (parameters, base loss, after_train, split / dup /
cull), `Cameras`, `get_viewmat`, `RGB2SH`, `Optimizers`, the camera optimizer in its "off" mode and the callback
types.  Nothing here is product code and nothing under fusionsense_b200/ imports it.

The reference's package comes from `baseline/_ref` (the task's sanctioned `pip install --no-deps --target baseline/_ref
/root/reference`, git-ignored); `install()` registers a bare `dn_splatter` package object for it so that the real
`dn_splatter/__init__.py` (which imports every dataparser and with them most of nerfstudio) does not run — the module
files themselves are executed as they are.
"""
from __future__ import annotations

import importlib
import sys
import types
from pathlib import Path

STUBS = Path(__file__).resolve().parent
ROOT = STUBS.parent.parent
REF = ROOT / "baseline" / "_ref"


class _Anything(types.ModuleType):
    """A module whose every attribute is a harmless placeholder (for imports the hot path never executes:
    matplotlib, open3d, natsort, ...)."""

    def __getattr__(self, name):
        if name.startswith("__"):
            raise AttributeError(name)
        obj = type(name, (), {"__init__": lambda self, *a, **k: None, "__call__": lambda self, *a, **k: None})
        setattr(self, name, obj)
        return obj


def reference_available() -> bool:
    return (REF / "dn_splatter" / "dn_model.py").exists()


def install() -> None:
    """Make `import nerfstudio / torchmetrics / gsplat / dn_splatter` resolve to the stubs, the shim and the reference."""
    if str(STUBS) not in sys.path:
        sys.path.insert(0, str(STUBS))
    for name in ("matplotlib", "matplotlib.pyplot", "matplotlib.cm", "open3d", "open3d.core", "natsort",
                 "nerfstudio.data.datasets", "nerfstudio.data.datasets.base_dataset", "nerfstudio.models.base_model",
                 "nerfstudio.process_data", "nerfstudio.process_data.process_data_utils", "nerfstudio.utils.colormaps"):
        if name not in sys.modules:
            try:
                importlib.import_module(name)
            except Exception:  # noqa: BLE001
                sys.modules[name] = _Anything(name)
    import fusionsense_b200

    fusionsense_b200.install_gsplat_shim(force=True)
    if "dn_splatter" not in sys.modules:
        pkg = types.ModuleType("dn_splatter")
        pkg.__path__ = [str(REF / "dn_splatter")]
        sys.modules["dn_splatter"] = pkg
