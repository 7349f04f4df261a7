"""Helpers to re-materialise the committed golden fixtures (tests/golden/*.npz)."""
from pathlib import Path

import numpy as np

GOLDEN = Path(__file__).resolve().parent / "golden"


def load_visual_hull_golden():
    """Outputs of the unmodified reference VisualHull() (oracle/make_golden_visual_hull.py) + its inputs."""
    z = np.load(GOLDEN / "visual_hull_201.npz")
    shape = tuple(int(v) for v in z["mask_shape"])
    masks = np.unpackbits(z["masks"], axis=-1)[..., : shape[-1]].astype(np.uint8) * 255
    return dict(c2w=z["c2w"], masks=masks.reshape(shape), fx=float(z["fx"]), points=z["points"],
                maxv=float(z["maxv"]), iso=float(z["iso"]))


def write_visual_hull_capture(tmpdir, g):
    from fusionsense_b200.synthetic import write_capture

    write_capture(str(tmpdir), masks=g["masks"], c2w=g["c2w"], fx=g["fx"], width=g["masks"].shape[2],
                  height=g["masks"].shape[1])
    return str(tmpdir)
