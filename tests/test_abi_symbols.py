"""CPU: libfsb200.so loads and exports every symbol include/fsb200.h declares (no compute calls: no GPU here),
the ctypes binding is generated from that header, and the product package has no route into oracle/."""
import ctypes
import re
import subprocess
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent


def test_header_symbols_are_exported():
    from fusionsense_b200 import _abi, _build

    assert _build.LIB_PATH.exists(), "run `python -c 'import __graft_entry__ as g; g.build()'` first"
    protos = _abi.parse_header()
    assert len(protos) >= 20
    cdll = ctypes.CDLL(str(_build.LIB_PATH))
    for name in protos:
        assert hasattr(cdll, name), f"{name} declared in include/fsb200.h but not exported"
    # and nothing un-declared leaks out of the library
    out = subprocess.run(["nm", "-D", "--defined-only", str(_build.LIB_PATH)], capture_output=True, text=True).stdout
    exported = set(re.findall(r"\b(fsb_\w+)\b", out))
    assert exported == set(protos), exported ^ set(protos)


def test_pure_host_entry_points_answer_without_a_gpu():
    from fusionsense_b200._abi import lib

    assert lib.fsb_abi_version() == 3
    assert lib.fsb_raster_supported_channels(3) == 3 and lib.fsb_raster_supported_channels(6) == 8
    assert lib.fsb_raster_supported_channels(33) == -1
    assert lib.fsb_isect_scan_workspace(5000) == 3 * 8
    assert lib.fsb_radix_sort_workspace(10000, 44) > 6 * 256 * 4
    assert lib.fsb_adam_max_tensors() == 8 and lib.fsb_vh_max_views() >= 9
    # argument validation happens before any CUDA call
    assert lib.fsb_radix_sort_pairs(-1, None, 44, None, None, None, None, None, 0, None, None) == 10001
    ws = lib.fsb_raster_workspace(0, 1, 0, 8)
    assert ws > 0
    assert lib.fsb_raster_fwd(1, 1, 7, 0, None, None, None, None, None, None, None, 16, 16, 16, 1, 1, None, None, 0, 1, ws,
                              None, None, None, None) == 10001  # D = 7 is not an instantiated channel count
    assert lib.fsb_raster_fwd(1, 1, 3, 0, None, None, None, None, None, None, None, 16, 16, 5, 1, 1, None, None, 0, 1, ws,
                              None, None, None, None) == 10001  # tile_size 5: not whole warps
    assert lib.fsb_raster_fwd(1, 1, 3, 0, None, None, None, None, None, None, None, 16, 16, 16, 1, 1, None, None, 0, None,
                              0, None, None, None, None) == 10001  # missing workspace


def test_product_package_never_imports_the_oracle():
    for py in (ROOT / "fusionsense_b200").rglob("*.py"):
        text = py.read_text()
        assert not re.search(r"^\s*(from|import)\s+oracle\b", text, re.M), f"{py} imports oracle/"
        assert "oracle." not in re.sub(r'""".*?"""', "", text, flags=re.S).replace("# ", ""), py


def test_ops_refuse_cpu_tensors():
    import torch

    from fusionsense_b200.gsplat import rasterization, rasterize_gaussians

    n = 4
    with pytest.raises(RuntimeError, match="CUDA"):
        rasterization(torch.zeros(n, 3), torch.ones(n, 4), torch.ones(n, 3), torch.ones(n), torch.ones(n, 3),
                      torch.eye(4)[None], torch.eye(3)[None], 32, 32, packed=False)
    with pytest.raises(RuntimeError, match="CUDA"):
        rasterize_gaussians(torch.zeros(n, 2), torch.ones(n), torch.ones(n, dtype=torch.int32), torch.ones(n, 3),
                            torch.ones(n, dtype=torch.int32), torch.ones(n, 3), torch.ones(n, 1), 32, 32, 16)


def test_gsplat_shim_resolves_reference_imports():
    import sys

    import fusionsense_b200

    saved = {k: v for k, v in sys.modules.items() if k == "gsplat" or k.startswith("gsplat.")}
    for k in saved:
        del sys.modules[k]
    try:
        fusionsense_b200.install_gsplat_shim()
        # the four imports of /root/reference/dn_splatter/dn_model.py:29-35
        from gsplat.rendering import rasterization  # noqa: F401
        from gsplat import rasterize_gaussians  # noqa: F401
        from gsplat.cuda_legacy._torch_impl import quat_to_rotmat
        from gsplat.cuda_legacy._wrapper import num_sh_bases
        import inspect

        assert [num_sh_bases(d) for d in range(5)] == [1, 4, 9, 16, 25]
        sig = inspect.signature(rasterization)
        assert list(sig.parameters)[:9] == ["means", "quats", "scales", "opacities", "colors", "viewmats", "Ks",
                                            "width", "height"]
        d = {k: v.default for k, v in sig.parameters.items() if v.default is not inspect._empty}
        assert d == dict(near_plane=0.01, far_plane=1e10, radius_clip=0.0, eps2d=0.3, sh_degree=None, packed=True,
                         tile_size=16, backgrounds=None, render_mode="RGB", sparse_grad=False, absgrad=False,
                         rasterize_mode="classic", channel_chunk=32)
        import torch

        q = torch.tensor([[2.0, 0.0, 0.0, 0.0], [0.0, 0.0, 0.0, 3.0]])
        R = quat_to_rotmat(q)
        assert torch.allclose(R[0], torch.eye(3)) and torch.allclose(R[1], torch.diag(torch.tensor([-1.0, -1.0, 1.0])))
    finally:
        for k in [k for k in sys.modules if k == "gsplat" or k.startswith("gsplat.")]:
            del sys.modules[k]
        sys.modules.update(saved)


def test_every_call_site_passes_the_declared_number_of_arguments():
    """ctypes checks types, not meaning: a call with a missing or extra argument would only fail on the GPU box.
    Walk the Python sources and compare every `lib.fsb_*(...)` call with the prototype parsed from the header."""
    import ast
    from pathlib import Path

    from fusionsense_b200._abi import parse_header

    protos = parse_header()
    root = Path(__file__).resolve().parent.parent
    files = list((root / "fusionsense_b200").rglob("*.py")) + [root / "bench.py", root / "__graft_entry__.py"]
    files += list((root / "tools").glob("*.py"))
    seen = 0
    for f in files:
        for node in ast.walk(ast.parse(f.read_text())):
            if not (isinstance(node, ast.Call) and isinstance(node.func, ast.Attribute)
                    and node.func.attr.startswith("fsb_") and isinstance(node.func.value, ast.Name)
                    and node.func.value.id == "lib"):
                continue
            name = node.func.attr
            assert name in protos, f"{f.name}:{node.lineno}: {name} is not declared in include/fsb200.h"
            if any(isinstance(a, ast.Starred) for a in node.args):
                continue
            seen += 1
            assert len(node.args) == len(protos[name][1]), (
                f"{f.name}:{node.lineno}: {name} called with {len(node.args)} arguments, header declares "
                f"{len(protos[name][1])}")
    assert seen > 40


def test_integration_md_stub_matches_the_header():
    """The hand-written ctypes stub shown in INTEGRATION.md §3 has as many argtypes as the header's prototype has
    parameters, pointer for pointer (a maintainer copying it must not get a drifted signature)."""
    import ctypes
    import re
    from pathlib import Path

    from fusionsense_b200 import _abi

    text = (Path(__file__).resolve().parent.parent / "INTEGRATION.md").read_text()
    block = text[text.index("lib.fsb_raster_fwd.argtypes"):]
    block = block[:block.index("\ndef raster_fwd")]
    expr = block.split("=", 1)[1].replace("\\\n", " ")
    P = ctypes.c_void_p
    argtypes = eval(expr, {"ctypes": ctypes, "P": P})  # noqa: S307 — our own document
    _, want, _ = _abi.parse_header()["fsb_raster_fwd"]
    assert len(argtypes) == len(want)
    assert [a is P for a in argtypes] == [w is ctypes.c_void_p for w in want]
    call = text[text.index("rc = lib.fsb_raster_fwd("):]
    call = call[:call.index("if rc:")]
    n_args = len(re.sub(r"\([^()]*\)", "", call[call.index("(") + 1:call.rindex(")")]).split(","))
    assert n_args == len(want)


def test_round2_host_mirrors_refuse_cpu_tensors_and_bad_arguments():
    """No CPU path behind the mirrors added in round 2 (KNN, density, level set, seed cloud, 8-bit targets), and the
    library refuses malformed arguments before any CUDA call."""
    import torch

    from fusionsense_b200 import knn, level_set, seed_points
    from fusionsense_b200._abi import lib
    from fusionsense_b200.compose import u8_to_unit_float

    x = torch.rand(50, 3)
    for call in (lambda: knn.knn_sk(x, x, 3),
                 lambda: knn.gaussian_density(x, torch.zeros(50, 4, dtype=torch.int64), x, x, torch.rand(50, 4),
                                              torch.rand(50, 1)),
                 lambda: level_set.level_crossings(x, [0.0, 0.0, 1.0], torch.zeros(50, 4, dtype=torch.int64), x, x,
                                                   torch.rand(50, 4), torch.rand(50, 1), [0.1, 0.3]),
                 lambda: seed_points.get_pointcloud(torch.rand(3, 4, 5), torch.rand(4, 5), torch.eye(4), 1.0, 1.0, 2.0, 2.0),
                 lambda: seed_points.voxel_down_sample(torch.rand(10, 6), 0.02),
                 lambda: u8_to_unit_float(torch.zeros(16, dtype=torch.uint8))):
        with pytest.raises(RuntimeError, match="CUDA"):
            call()
    E = 10001  # FSB_E_ARG
    assert lib.fsb_knn_cells(10, None, 0, None, None, None, None, None) == E        # g = 0
    assert lib.fsb_knn_cells(10, None, 321, None, None, None, None, None) == E      # more cells per axis than the cap
    assert lib.fsb_knn_query(10, None, None, 4, None, None, None, 34, 0, 8, None, None, None, None, None) == E  # K > 33
    assert lib.fsb_knn_query(10, None, None, 4, None, None, None, 5, 5, 8, None, None, None, None, None) == E   # drops all
    assert lib.fsb_gaussian_density(-1, None, 4, None, None, None, None, None, 0.0, None, None) == E
    assert lib.fsb_gaussian_density(0, None, 4, None, None, None, None, None, 0.0, None, None) == 0            # empty: no-op
    assert lib.fsb_level_crossings(10, None, None, 16, None, None, None, None, None, None, 5, None, None, None, None,
                                   None, None) == E                                                          # 5 levels
    assert lib.fsb_voxel_keys(10, None, 6, 0.0, None, None, None, None, None, 0, None) == E                    # voxel <= 0
    assert lib.fsb_voxel_mean(10, None, None, None, None, None, 6, 10, None, None) == E                        # width > 9
    assert lib.fsb_u8_to_unit_float(-1, None, None, 1, None) == E and lib.fsb_u8_to_unit_float(0, None, None, 1, None) == 0
    assert lib.fsb_backproject_emit(0, 4, None, None, None, None, 1.0, 1.0, 0.0, 0.0, 0.0, 1.0, None, None, None) == E
