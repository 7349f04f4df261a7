"""pin the CPU oracle (oracle/gsplat_ref.py) against a REAL gsplat==1.0.0, on the first box
that has one (never the case so far: gsplat is not vendored in the reference tree, there is no wheel and no network;
the oracle therefore says "parity unpinned").

  python tools/check_vs_gsplat.py [--n 50000] [--out gpurun_out/check_vs_gsplat.json]

What it does when `import gsplat` resolves to something that is not this repository's shim:
  1. records the facts the oracle was written from: version, signatures of rasterization / rasterize_gaussians,
     num_sh_bases source, the constants in the CUDA sources (0.999, 1/255, 1e-4, 0.3, 1.3);
  2. runs cfg1 (50k Gaussians, 640x480, SH degree 3, RGB+ED + legacy normals pass) through gsplat's CUDA kernels and
     through the oracle on the CPU with the same inputs, and reports max / mean relative error of every output and
     gradient plus the fraction of differing sort keys;
  3. does the same for this repository's kernels, so the three-way agreement is on one page.
Any disagreement beyond fp32 noise means Appendix A is wrong: fix the oracle, not the tolerance."""
import argparse
import inspect
import json
import re
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))


def find_real_gsplat():
    for extra in (ROOT / "baseline" / "_ref", None):
        if extra is not None and extra.is_dir():
            sys.path.insert(0, str(extra))
        try:
            import importlib

            for m in [k for k in sys.modules if k == "gsplat" or k.startswith("gsplat.")]:
                del sys.modules[m]
            g = importlib.import_module("gsplat")
            f = str(Path(getattr(g, "__file__", "")).resolve())
            if str(ROOT / "fusionsense_b200") in f or str(ROOT / "shim") in f:
                continue
            return g
        except ImportError:
            continue
    return None


def rel(a, b):
    import torch

    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    s = b.abs().max().clamp(min=1e-30)
    e = (a - b).abs() / s
    return {"max": float(e.max()), "mean": float(e.mean()), "frac_over_1e-4": float((e > 1e-4).double().mean())}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--n", type=int, default=50_000)
    ap.add_argument("--out", default=str(ROOT / "gpurun_out" / "check_vs_gsplat.json"))
    args = ap.parse_args()
    g = find_real_gsplat()
    out = {"gsplat": None}
    if g is None:
        out["status"] = "no real gsplat importable (only this repository's shim): the oracle stays unpinned"
        print(json.dumps(out))
        return 0
    import torch

    from fusionsense_b200.synthetic import make_scene
    from oracle import gsplat_ref as oracle

    out["gsplat"] = {"version": getattr(g, "__version__", "?"), "file": g.__file__,
                     "rasterization": str(inspect.signature(g.rendering.rasterization)),
                     "rasterize_gaussians": str(inspect.signature(g.rasterize_gaussians))}
    try:
        import gsplat.cuda_legacy._wrapper as w

        out["gsplat"]["num_sh_bases"] = inspect.getsource(w.num_sh_bases)
    except Exception as exc:  # noqa: BLE001
        out["gsplat"]["num_sh_bases"] = f"unavailable: {exc}"
    consts = {}
    for src in Path(g.__file__).parent.glob("cuda/csrc/*.cu*"):
        text = src.read_text(errors="ignore")
        for c in ("0.999f", "1.f / 255.f", "1e-4f", "0.3f", "1.3f"):
            if c in text:
                consts.setdefault(c, []).append(src.name)
    out["gsplat"]["constants_found"] = consts

    W, H = 640, 480
    sc = make_scene(args.n, W, H, n_views=1, cfg_id=1)
    colors = torch.cat((sc.features_dc[:, None, :], sc.features_rest), dim=1)

    def run(mod, dev):
        P = lambda t: t.clone().to(dev).requires_grad_(True)  # noqa: E731
        means, quats, scales, opac, cols = P(sc.means), P(sc.quats), P(sc.scales), P(sc.opacities), P(colors)
        render, alpha, info = mod.rasterization(
            means=means, quats=quats / quats.norm(dim=-1, keepdim=True), scales=torch.exp(scales),
            opacities=torch.sigmoid(opac).squeeze(-1), colors=cols, viewmats=sc.viewmats[:1].to(dev),
            Ks=sc.Ks[:1].to(dev), width=W, height=H, tile_size=16, packed=False, near_plane=0.01, far_plane=1e10,
            render_mode="RGB+ED", sh_degree=3, sparse_grad=False, absgrad=False, rasterize_mode="classic")
        nrm = torch.nn.functional.normalize(torch.randn(args.n, 3, generator=torch.Generator().manual_seed(2)), dim=-1).to(dev)
        nim = mod.rasterize_gaussians(info["means2d"][0].detach(), info["depths"][0], info["radii"][0],
                                      info["conics"][0], info["tiles_per_gauss"][0], nrm, torch.sigmoid(opac), H, W, 16)
        gen = torch.Generator().manual_seed(3)
        (render * torch.randn(render.shape, generator=gen).to(dev)).sum().backward(retain_graph=True)
        (nim * torch.randn(nim.shape, generator=gen).to(dev)).sum().backward()
        grads = {k: v.grad for k, v in dict(means=means, quats=quats, scales=scales, opacities=opac, colors=cols).items()}
        return dict(render=render, alpha=alpha, normals=nim, radii=info["radii"], isect_ids=info["isect_ids"],
                    flatten_ids=info["flatten_ids"], **{f"grad.{k}": v for k, v in grads.items()})

    res_ref = run(oracle, "cpu")
    res_gs = run(g, "cuda")
    import fusionsense_b200.gsplat as ours

    res_us = run(ours, "cuda")
    for name, res in (("gsplat_vs_oracle", res_gs), ("ours_vs_oracle", res_us)):
        rep = {}
        for k in res_ref:
            if k in ("radii", "isect_ids", "flatten_ids"):
                a, b = res[k].cpu(), res_ref[k]
                rep[k] = {"equal": bool(a.shape == b.shape and torch.equal(a, b)),
                          "n": int(b.numel()), "n_other": int(a.numel())}
            else:
                rep[k] = rel(res[k], res_ref[k])
        out[name] = rep
    Path(args.out).parent.mkdir(parents=True, exist_ok=True)
    Path(args.out).write_text(json.dumps(out, indent=1))
    print(json.dumps({k: v for k, v in out.items() if k != "gsplat"}, indent=1))
    return 0


if __name__ == "__main__":
    sys.exit(main())
