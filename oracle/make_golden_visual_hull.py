"""Generate tests/golden/visual_hull_201.npz by running the UNMODIFIED reference function
/root/reference/utils/VisualHull.py::VisualHull (lines 87-200) in this container.

Run from the repo root:  python -m oracle.make_golden_visual_hull
Needs /root/reference (absent on the GPU box, which only reads the committed .npz).
open3d / matplotlib are not installed here, so they are stubbed in sys.modules (the stub records the point
array handed to o3d.utility.Vector3dVector); PIL.Image.fromarray is shimmed because utils/readCam.py:50 builds
an int8 array that Pillow >= 12 rejects.  None of that touches the carving arithmetic.
"""
import hashlib
import sys
import tempfile
import types
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
REF = Path("/root/reference")


def _stub_modules(captured):
    o3d = types.ModuleType("open3d")

    class _PC:
        points = None

    o3d.geometry = types.SimpleNamespace(PointCloud=_PC)

    def _v3(a):
        captured["points"] = np.array(a, dtype=np.float64, copy=True)
        return a

    o3d.utility = types.SimpleNamespace(Vector3dVector=_v3)
    o3d.io = types.SimpleNamespace(write_point_cloud=lambda *a, **k: True)
    sys.modules["open3d"] = o3d
    mpl = types.ModuleType("matplotlib")
    plt = types.ModuleType("matplotlib.pyplot")

    class _Ax:
        def __getattr__(self, k):
            return lambda *a, **kw: None

    class _Fig:
        def add_subplot(self, *a, **k):
            return _Ax()

    plt.figure = lambda *a, **k: _Fig()
    plt.savefig = lambda *a, **k: None
    mpl.pyplot = plt
    sys.modules["matplotlib"] = mpl
    sys.modules["matplotlib.pyplot"] = plt
    from PIL import Image

    _orig = Image.fromarray

    def _fromarray(arr, mode=None):
        return _orig(np.asarray(arr).astype(np.uint8))

    Image.fromarray = _fromarray


def main():
    sys.path.insert(0, str(ROOT))
    from fusionsense_b200.synthetic import write_capture

    captured = {}
    _stub_modules(captured)
    sys.path.insert(0, str(REF))
    import io
    from contextlib import redirect_stdout

    from utils.VisualHull import VisualHull  # the reference, unmodified

    with tempfile.TemporaryDirectory() as d:
        c2w, masks = write_capture(d, n_views=9)
        buf = io.StringIO()
        with redirect_stdout(buf):
            VisualHull(d, d, error=5)
        log = buf.getvalue()
    pts = captured["points"]
    maxv = float([l for l in log.splitlines() if l.startswith("max number of votes:")][0].split(":")[1])
    iso = float([l for l in log.splitlines() if l.startswith("threshold for marching cube:")][0].split(":")[1])
    out = ROOT / "tests" / "golden" / "visual_hull_201.npz"
    np.savez_compressed(out, c2w=c2w, masks=np.packbits(masks > 0, axis=-1), mask_shape=np.array(masks.shape),
                        fx=600.0, points=pts, maxv=maxv, iso=iso,
                        points_sha256=np.frombuffer(hashlib.sha256(pts.tobytes()).digest(), dtype=np.uint8),
                        numpy_version=np.array(np.__version__))
    print(f"wrote {out}: {pts.shape[0]} occupied voxels, maxv={maxv}, iso={iso}, {out.stat().st_size} bytes")


if __name__ == "__main__":
    main()
