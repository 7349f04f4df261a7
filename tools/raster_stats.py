"""Workload statistics of the compositing kernels on the bench scene, computed on the CPU from the oracle's
projection and binning (no GPU needed): how long the per-tile lists are, how many pixels of a tile / of a warp
footprint an entry really blends into, where pixels stop.  These are the quantities the kernel design in
csrc/raster.cu trades against each other (segment length, footprint shape, per-entry warp reduction).

  python tools/raster_stats.py [cfg2|cfg1] [view] [out.json]
"""
import json
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import numpy as np
import torch

from fusionsense_b200.synthetic import make_scene
from oracle import gsplat_ref as ref

CFGS = {
    "cfg1": dict(n=50_000, w=640, h=480, kind="random", views=3, cfg_id=1),
    "cfg2": dict(n=300_000, w=640, h=480, kind="bunny", views=9, cfg_id=2),
}
ALPHA_MIN, ALPHA_MAX, T_MIN = 1.0 / 255.0, 0.999, 1e-4
SEG = 512
FOOT = [(8, 4), (16, 2), (4, 8)]  # (width, height) of a warp's pixel footprint inside the 16x16 tile


def rect_reach(gx, gy, opac, a, b, c, x0, x1, y0, y1):
    """The test of csrc/raster.cu::strip_mask on one rectangle of pixel centres [x0, x1] x [y0, y1] (float32, the
    kernel's arithmetic): can alpha reach 1/255 anywhere in it?  min over the rectangle of
    q(d) = a dx^2 + 2 b dx dy + c dy^2 against 2 ln(255 opacity), inflated; doubtful inputs answer yes."""
    f = np.float32
    gx, gy, opac, a, b, c = (np.asarray(v, dtype=f) for v in (gx, gy, opac, a, b, c))
    tau = np.log(f(255.0) * opac).astype(f)
    never = tau + f(2e-3) < 0
    det = a * c - b * b
    doubtful = ~(det > 0) | ~(a > 0) | ~(tau < f(1e30))
    thr = f(2.0) * tau * f(1.0002) + f(1e-2)
    X0, X1, Y0, Y1 = f(x0) - gx, f(x1) - gx, f(y0) - gy, f(y1) - gy
    inside = (X0 <= 0) & (X1 >= 0) & (Y0 <= 0) & (Y1 >= 0)
    with np.errstate(all="ignore"):
        nb_c, nb_a = -b / c, -b / a
        t = np.minimum(np.maximum(nb_c * X0, Y0), Y1); q = a * X0 * X0 + f(2) * b * X0 * t + c * t * t
        t = np.minimum(np.maximum(nb_c * X1, Y0), Y1); q = np.minimum(q, a * X1 * X1 + f(2) * b * X1 * t + c * t * t)
        t = np.minimum(np.maximum(nb_a * Y0, X0), X1); q = np.minimum(q, a * t * t + f(2) * b * t * Y0 + c * Y0 * Y0)
        t = np.minimum(np.maximum(nb_a * Y1, X0), X1); q = np.minimum(q, a * t * t + f(2) * b * t * Y1 + c * Y1 * Y1)
    q = np.where(inside, f(0), q)
    return ~never & (doubtful | ~(q > thr))


def main():
    name = sys.argv[1] if len(sys.argv) > 1 else "cfg2"
    view = int(sys.argv[2]) if len(sys.argv) > 2 else 0
    out_path = sys.argv[3] if len(sys.argv) > 3 else None
    c = CFGS[name]
    sc = make_scene(c["n"], c["w"], c["h"], n_views=c["views"], cfg_id=c["cfg_id"], kind=c["kind"])
    W, H, ts = c["w"], c["h"], 16
    tw, th = (W + ts - 1) // ts, (H + ts - 1) // ts
    with torch.no_grad():
        q = sc.quats / sc.quats.norm(dim=-1, keepdim=True)
        radii, m2, depths, conics, _ = ref.fully_fused_projection(
            sc.means, q, torch.exp(sc.scales), sc.viewmats[view:view + 1], sc.Ks[view:view + 1], W, H)
        _, ids, flat = ref.isect_tiles(m2, radii, depths, ts, tw, th)
        offs = ref.isect_offset_encode(ids, 1, tw, th).reshape(-1).tolist() + [flat.numel()]
        op = torch.sigmoid(sc.opacities[:, 0])
    m2, conics = m2[0].numpy().astype(np.float64), conics[0].numpy().astype(np.float64)
    op = op.numpy().astype(np.float64)
    flat = flat.numpy()
    lens = np.diff(np.array(offs))
    n_vis = int((radii > 0).sum())

    hist_tile = np.zeros(257, dtype=np.int64)                     # blended pixels per (tile, entry)
    hist_warp = {f: np.zeros(33, dtype=np.int64) for f in FOOT}   # blended lanes per (warp, entry), >= 1 only
    reach_warp = {f: 0 for f in FOOT}                             # (warp, entry) pairs with any alpha-test pass
    stop_pixels = 0
    stop_segment_hist = np.zeros(64, dtype=np.int64)              # segment index in which a pixel stops
    blended = visited = 0
    segs_needed = 0                                               # segments up to the tile's deepest last id
    geo_miss = 0                                                  # (tile, entry): alpha test fails on every pixel
    pruned_lens = np.zeros(tw * th, dtype=np.int64)               # list lengths without the geometric misses
    rect_kept = rect_false_prunes = 0                             # the kernel's rectangle test at tile granularity
    for lin in range(tw * th):
        s, e = offs[lin], offs[lin + 1]
        if e <= s:
            continue
        ty, tx = divmod(lin, tw)
        g = flat[s:e]
        py, px = np.meshgrid(np.arange(ty * ts, ty * ts + ts) + 0.5, np.arange(tx * ts, tx * ts + ts) + 0.5,
                             indexing="ij")
        inside = ((py < H) & (px < W)).reshape(-1)
        dx = m2[g, 0][None] - px.reshape(-1, 1)
        dy = m2[g, 1][None] - py.reshape(-1, 1)
        sigma = 0.5 * (conics[g, 0] * dx * dx + conics[g, 2] * dy * dy) + conics[g, 1] * dx * dy
        alpha = np.minimum(op[g][None] * np.exp(-sigma), ALPHA_MAX)
        valid = (sigma >= 0) & (alpha >= ALPHA_MIN) & inside[:, None]
        a_eff = np.where(valid, alpha, 0.0)
        T_incl = np.cumprod(1.0 - a_eff, axis=1)
        stop = valid & (T_incl <= T_MIN)
        dead = np.cumsum(stop, axis=1) > 0
        contrib = valid & ~dead                                    # [256, G]
        blended += int(contrib.sum())
        any_c = contrib.any(axis=1)
        last = np.where(any_c, contrib.shape[1] - 1 - np.argmax(contrib[:, ::-1], axis=1), -1)
        visited += int((last + 1).sum())
        segs_needed += int(last.max() // SEG + 1) if last.max() >= 0 else 0
        stopped = dead[:, -1]
        stop_pixels += int(stopped.sum())
        first_stop = np.argmax(dead, axis=1)
        np.add.at(stop_segment_hist, np.minimum(first_stop[stopped] // SEG, 63), 1)
        np.add.at(hist_tile, contrib.sum(axis=0), 1)
        hit = valid.any(axis=0)
        geo_miss += int((~hit).sum())
        pruned_lens[lin] = int(hit.sum())
        keep = rect_reach(m2[g, 0], m2[g, 1], op[g], conics[g, 0], conics[g, 1], conics[g, 2],
                          tx * ts + 0.5, min(tx * ts + ts, W) - 0.5, ty * ts + 0.5, min(ty * ts + ts, H) - 0.5)
        rect_kept += int(keep.sum())
        rect_false_prunes += int((hit & ~keep).sum())
        grid_c = contrib.reshape(ts, ts, -1)
        grid_v = valid.reshape(ts, ts, -1)
        for (fw, fh) in FOOT:
            cc = grid_c.reshape(ts // fh, fh, ts // fw, fw, -1).sum(axis=(1, 3)).reshape(-1)
            vv = grid_v.reshape(ts // fh, fh, ts // fw, fw, -1).any(axis=(1, 3)).reshape(-1)
            np.add.at(hist_warp[(fw, fh)], cc[cc > 0], 1)
            reach_warp[(fw, fh)] += int(vv.sum())

    def summarize(h):
        n = int(h[1:].sum())
        k = np.arange(len(h))
        cum = np.cumsum(h[1:]) / max(n, 1)
        return {"pairs": n, "mean_lanes": float((h * k).sum() / max(n, 1)),
                "frac_le_1": float(cum[0]), "frac_le_2": float(cum[1]), "frac_le_4": float(cum[3]),
                "frac_le_8": float(cum[7]), "frac_le_16": float(cum[15])}

    n_tiles_used = int((lens > 0).sum())
    out = {
        "config": name, "view": view, "n_visible": n_vis, "n_isects": int(lens.sum()), "tiles": tw * th,
        "list_len": {"median": float(np.median(lens)), "p90": float(np.percentile(lens, 90)),
                     "p99": float(np.percentile(lens, 99)), "max": int(lens.max()),
                     "tiles_over_one_segment": int((lens > SEG).sum()),
                     "segments_total": int(np.maximum(1, -(-lens // SEG)).sum()),
                     "segments_up_to_deepest_last_id": segs_needed,
                     "isects_in_multi_segment_tiles": int(lens[lens > SEG].sum())},
        "pairs": {"blended": blended, "visited_by_a_pixel_walk": visited,
                  "blended_per_tile_entry_mean": blended / max(int(lens.sum()), 1),
                  "tile_entries_with_no_blended_pixel": int(hist_tile[0]),
                  "tile_entries_failing_the_alpha_test_on_every_pixel": geo_miss},
        "list_len_without_geometric_misses": {
            "n_isects": int(pruned_lens.sum()), "median": float(np.median(pruned_lens)),
            "p99": float(np.percentile(pruned_lens, 99)), "max": int(pruned_lens.max()),
            "tiles_over_one_segment": int((pruned_lens > SEG).sum()),
            "segments_total": int(np.maximum(1, -(-pruned_lens // SEG)).sum())},
        "tile_rectangle_test": {"kept": rect_kept, "pruned": int(lens.sum()) - rect_kept,
                                "entries_pruned_that_touch_a_pixel": rect_false_prunes},
        "stop": {"pixels_that_stop": stop_pixels, "of_pixels": W * H,
                 "stop_segment_hist": {str(i): int(v) for i, v in enumerate(stop_segment_hist) if v}},
        "per_warp_entry": {f"{fw}x{fh}": {**summarize(hist_warp[(fw, fh)]), "reached_pairs": reach_warp[(fw, fh)]}
                           for (fw, fh) in FOOT},
        "tiles_nonempty": n_tiles_used,
    }
    txt = json.dumps(out, indent=1)
    print(txt)
    if out_path:
        Path(out_path).write_text(txt)


if __name__ == "__main__":
    main()
