"""Torch-facing wrappers of the libfsb200 C-ABI: raw stage calls and the autograd Functions.

PyTorch is plumbing here (device memory, streams, autograd graph); all arithmetic of the render
path runs in the hand-written sm_100a kernels.  There is no CPU path: every function requires CUDA
tensors and raises otherwise.
"""
from __future__ import annotations

import ctypes
import os
import math
from typing import Optional, Tuple

import torch
from torch import Tensor

from ._abi import check, lib, ptr


def _stream() -> int:
    """Raw cudaStream_t of torch's current stream on the current device (the cheap private accessor when this
    torch has it: torch.cuda.current_stream() builds a Stream object, ~10x slower, and runs ~20 times per step)."""
    try:
        return torch._C._cuda_getCurrentRawStream(torch.cuda.current_device())
    except AttributeError:  # pragma: no cover
        return torch.cuda.current_stream().cuda_stream


# FSB_NVTX=1: every stage call wrapped by kernel_timer.start / stop (projection, emit, sort, compositing forward /
# backward, SSIM, Adam) is also an NVTX range "fsb.<stage>", so nsys / ncu --nvtx timelines name the stages
# (SURVEY.md §5 asked for ranges; off by default: a push / pop pair costs ~1 us of host time per call)
NVTX = __import__("os").environ.get("FSB_NVTX", "0") == "1"


class KernelTimer:
    """Optional CUDA-event timing of individual ABI calls on the launching stream (bench.py's roofline leg).

    Disabled by default (zero overhead); `with ops.kernel_timer.collect(): ...` records (start, end) events
    around the calls wrapped with `_timed`, `summary()` returns mean milliseconds per name after a sync.
    """

    def __init__(self):
        self.enabled = False
        self.events = {}
        self.pad_cycles = 0

    def collect(self, pad_cycles: int = 0):
        """`pad_cycles` > 0: a spin kernel of that many SM cycles is queued before every start event, so the host
        has recorded the event and launched the timed kernels before the device gets to them — otherwise, in an
        eager (host-bound) step, the interval also counts the time the idle device waits for the launch."""
        timer = self

        class _Ctx:
            def __enter__(self_inner):
                timer.enabled, timer.events, timer.pad_cycles = True, {}, int(pad_cycles)
                return timer

            def __exit__(self_inner, *exc):
                timer.enabled = False
                return False

        return _Ctx()

    def start(self, name):
        if NVTX:
            torch.cuda.nvtx.range_push(f"fsb.{name}")
        if not self.enabled:
            return None
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        if self.pad_cycles:
            torch.cuda._sleep(self.pad_cycles)
        s.record(torch.cuda.current_stream())
        self.events.setdefault(name, []).append((s, e))
        return e

    @staticmethod
    def stop(e):
        if NVTX:
            torch.cuda.nvtx.range_pop()
        if e is not None:
            e.record(torch.cuda.current_stream())

    def summary(self):
        torch.cuda.synchronize()
        return {k: (sum(s.elapsed_time(e) for s, e in v) / len(v), len(v)) for k, v in self.events.items()}


kernel_timer = KernelTimer()


class StaticCapacity:
    """Static-capacity mode of the intersection pipeline (include/fsb200.h): list buffers are allocated for
    `capacity` entries, the true count stays on the device, nothing is read back to the host, and the launch
    sequence of a whole render + backward is fixed, so a CUDA graph can capture it.  `overflow` is a device int32[1]
    that fsb_isect_emit raises when a count exceeds the capacity (the step's results are then invalid)."""

    def __init__(self, capacity: int, overflow: Tensor):
        assert capacity > 0 and overflow.dtype == torch.int32 and overflow.is_cuda
        self.capacity, self.overflow = int(capacity), overflow
        self.counts = []  # device int64[1] tensors of the true counts seen in this scope (for reporting)


_static: Optional[StaticCapacity] = None


def static_mode() -> Optional[StaticCapacity]:
    return _static


class static_capacity:
    """`with ops.static_capacity(capacity, overflow_flag): ...` — see StaticCapacity."""

    def __init__(self, capacity: int, overflow: Tensor):
        self.state = StaticCapacity(capacity, overflow)

    def __enter__(self):
        global _static
        self.prev, _static = _static, self.state
        return self.state

    def __exit__(self, *exc):
        global _static
        _static = self.prev
        return False


def _req_cuda(*ts):
    for t in ts:
        if t is not None and not t.is_cuda:
            raise RuntimeError("fusionsense_b200 ops need CUDA tensors (there is no CPU fallback)")


def _f32c(t: Optional[Tensor]) -> Optional[Tensor]:
    if t is None:
        return None
    if t.dtype != torch.float32:
        t = t.float()
    return t.contiguous()


# ---------------------------------------------------------------------------------------------
# raw stage calls
# ---------------------------------------------------------------------------------------------
def project_sh_fwd(means, quats, scales, viewmats, Ks, width, height, eps2d, near_plane, far_plane, radius_clip,
                   tile_size, sh_degree, coeffs, campos, color_stride, depth_channel, calc_comp, legacy_extra=None):
    _req_cuda(means, quats, scales, viewmats, Ks)
    C, N = viewmats.shape[0], means.shape[0]
    dev = means.device
    tile_w, tile_h = math.ceil(width / tile_size), math.ceil(height / tile_size)
    radii = torch.empty((C, N), dtype=torch.int32, device=dev)
    means2d = torch.empty((C, N, 2), dtype=torch.float32, device=dev)
    depths = torch.empty((C, N), dtype=torch.float32, device=dev)
    conics = torch.empty((C, N, 3), dtype=torch.float32, device=dev)
    comps = torch.empty((C, N), dtype=torch.float32, device=dev) if calc_comp else None
    colors = torch.empty((C, N, color_stride), dtype=torch.float32, device=dev) if color_stride > 0 else None
    tiles = torch.empty((C, N), dtype=torch.int32, device=dev)
    K = coeffs.shape[1] if coeffs is not None else 0
    ev = kernel_timer.start("project_sh_fwd")
    check(lib.fsb_project_sh_fwd(C, N, ptr(means), ptr(quats), ptr(scales), ptr(viewmats), ptr(Ks), width, height,
                                 eps2d, near_plane, far_plane, radius_clip, tile_size, tile_w, tile_h,
                                 -1 if sh_degree is None else sh_degree, K, ptr(coeffs), ptr(campos), color_stride,
                                 depth_channel, ptr(radii), ptr(means2d), ptr(depths), ptr(conics), ptr(comps),
                                 ptr(colors), ptr(tiles), ptr(legacy_extra), _stream()), "fsb_project_sh_fwd")
    kernel_timer.stop(ev)
    return radii, means2d, depths, conics, comps, colors, tiles


def isect_count(means2d: Tensor, radii: Tensor, tile_size: int, tile_w: int, tile_h: int, legacy_bbox: bool):
    _req_cuda(means2d, radii)
    M = radii.numel()
    tiles = torch.empty(radii.shape, dtype=torch.int32, device=radii.device)
    check(lib.fsb_isect_count(M, ptr(means2d), ptr(radii), tile_size, tile_w, tile_h, int(legacy_bbox), ptr(tiles),
                              _stream()), "fsb_isect_count")
    return tiles


def isect_count_reach(means2d: Tensor, radii: Tensor, conics: Tensor, opacities: Tensor, tile_size: int, tile_w: int,
                      tile_h: int, legacy_bbox: bool, depths: Optional[Tensor] = None):
    """csrc/isect_reach.cu: tiles of each Gaussian's bounding box on which it can pass the alpha test.
    means2d [C,N,2], radii [C,N], conics [C,N,3], opacities [C,N] -> counts [C,N] int32.
    `depths` [C,N]: also write the (depth bits, index) pairs of the two-level binning's depth sort
    (counts.depth_keys int64 [C*N], counts.depth_vals int32 [C*N])."""
    _req_cuda(means2d, radii, conics, opacities)
    C, N = radii.shape
    dev = radii.device
    counts = torch.empty(radii.shape, dtype=torch.int32, device=dev)
    hit_masks = torch.empty(radii.shape, dtype=torch.int64, device=dev)  # read back by isect_emit(reach=...)
    depth_keys = depth_vals = None
    if depths is not None:
        _req_cuda(depths)
        depth_keys = torch.empty((C * N,), dtype=torch.int64, device=dev)
        depth_vals = torch.empty((C * N,), dtype=torch.int32, device=dev)
    check(lib.fsb_isect_count_reach(C, N, ptr(means2d), ptr(radii), ptr(conics), ptr(opacities), tile_size, tile_w,
                                    tile_h, int(legacy_bbox), ptr(counts), ptr(hit_masks), ptr(depths), ptr(depth_keys),
                                    ptr(depth_vals), _stream()),
          "fsb_isect_count_reach")
    counts.hit_masks = hit_masks
    counts.depth_keys, counts.depth_vals = depth_keys, depth_vals
    return counts


def isect_scan(counts: Tensor, totals: Optional[Tensor] = None, read_back: bool = True, perm: Optional[Tensor] = None):
    """Exclusive int64 offsets of `counts` and the total (one small D2H read, like gsplat's).
    `perm` (int32 [M]): the scan runs over counts[perm[i]] (two-level binning: the Gaussians in depth order).

    `totals`: optional int64 device tensor whose element 0 receives the total; the whole tensor comes back in the
    same read (rasterization() keeps the legacy-bbox mismatch counter in element 1), and the return value is then
    (offsets, list_of_ints)."""
    M = counts.numel()
    dev = counts.device
    offsets = torch.empty((M,), dtype=torch.int64, device=dev)
    total = totals if totals is not None else torch.empty((1,), dtype=torch.int64, device=dev)
    ws_bytes = lib.fsb_isect_scan_workspace(M)
    ws = torch.empty((ws_bytes,), dtype=torch.uint8, device=dev)
    if perm is None:
        check(lib.fsb_isect_scan(M, ptr(counts), ptr(offsets), ptr(total), ptr(ws), ws_bytes, _stream()),
              "fsb_isect_scan")
    else:
        check(lib.fsb_isect_scan_perm(M, ptr(counts), ptr(perm), ptr(offsets), ptr(total), ptr(ws), ws_bytes, _stream()),
              "fsb_isect_scan_perm")
    if not read_back:
        return offsets, None
    if totals is not None:
        return offsets, [int(v) for v in total.tolist()]
    return offsets, int(total.item())


def tile_bits_for(n_tiles: int) -> int:
    return int(n_tiles).bit_length()


def sort_end_bit(n_tiles: int, C: int) -> int:
    """Key bits that can differ: 32 depth bits + tile bits (+ camera bits when there is more than one camera; gsplat
    reserves floor(log2 C) + 1 bits even for C = 1, whose value is always 0).  One bit less is one radix pass less at
    1080p: 45 bits sort in five 9-bit passes, 46 in six 8-bit ones (csrc/radix_sort.cu)."""
    return 32 + tile_bits_for(n_tiles) + (tile_bits_for(C) if C > 1 else 0)


def isect_emit(means2d, radii, depths, offsets, n_isects, C, N, tile_size, tile_w, tile_h, legacy_bbox, n_dev=None,
               overflow=None, reach=None, hit_masks=None, perm=None, packed=False):
    """`n_dev` (device int64[1]) selects static-capacity mode: `n_isects` is then the capacity of the buffers.
    `reach` = (conics [C,N,3], opacities [C,N]): emit only the reached tiles (`offsets` must then be the scan of
    `isect_count_reach`).  `perm` / `packed` (reach only): two-level binning, see include/fsb200.h; with `packed` the
    second return value is None."""
    dev = means2d.device
    ids = torch.empty((n_isects,), dtype=torch.int64, device=dev)
    flat = None if packed else torch.empty((n_isects,), dtype=torch.int32, device=dev)
    assert reach is not None or (perm is None and not packed)
    tb = tile_bits_for(tile_w * tile_h)
    ev = kernel_timer.start("isect_emit")
    if reach is None:
        check(lib.fsb_isect_emit(C, N, ptr(means2d), ptr(radii), ptr(depths), ptr(offsets), tile_size, tile_w, tile_h,
                                 tb, int(legacy_bbox), ptr(n_dev), n_isects, ptr(overflow), ptr(ids), ptr(flat),
                                 _stream()), "fsb_isect_emit")
    else:
        check(lib.fsb_isect_emit_reach(C, N, ptr(means2d), ptr(radii), ptr(depths), ptr(reach[0]), ptr(reach[1]),
                                       ptr(offsets), tile_size, tile_w, tile_h, tb, int(legacy_bbox), ptr(n_dev),
                                       n_isects, ptr(overflow), ptr(ids), ptr(flat), ptr(hit_masks), ptr(perm), int(packed),
                                       _stream()),
              "fsb_isect_emit_reach")
    kernel_timer.stop(ev)
    return ids, flat


def radix_sort_pairs(keys: Tensor, vals: Tensor, end_bit: int, n_dev: Optional[Tensor] = None) -> Tuple[Tensor, Tensor]:
    """Stable sort of (int64 keys, int32 vals) on key bits [0, end_bit). Inputs are clobbered.
    `n_dev`: static-capacity mode, only the first min(*n_dev, len) pairs are sorted (the rest is untouched)."""
    _req_cuda(keys, vals)
    n = keys.numel()
    if n == 0:
        return keys, vals
    dev = keys.device
    keys_b = torch.empty_like(keys)
    vals_b = torch.empty_like(vals)
    ws_bytes = lib.fsb_radix_sort_workspace(n, end_bit)
    ws = torch.empty((ws_bytes,), dtype=torch.uint8, device=dev)
    in_b = ctypes.c_int(0)
    ev = kernel_timer.start("radix_sort")
    check(lib.fsb_radix_sort_pairs(n, ptr(n_dev), end_bit, ptr(keys), ptr(vals), ptr(keys_b), ptr(vals_b), ptr(ws), ws_bytes,
                                   ctypes.addressof(in_b), _stream()), "fsb_radix_sort_pairs")
    kernel_timer.stop(ev)
    return (keys_b, vals_b) if in_b.value else (keys, vals)


def radix_sort_keys(keys: Tensor, begin_bit: int, end_bit: int, n_dev: Optional[Tensor] = None,
                    want_low32: bool = True) -> Tuple[Tensor, Optional[Tensor]]:
    """Stable sort of int64 keys on bits [begin_bit, end_bit); the input is clobbered.  Returns (sorted keys, int32 low
    words of the sorted keys).  `n_dev` as in radix_sort_pairs."""
    _req_cuda(keys)
    n = keys.numel()
    low = torch.empty((n,), dtype=torch.int32, device=keys.device) if want_low32 else None
    if n == 0:
        return keys, low
    keys_b = torch.empty_like(keys)
    ws_bytes = lib.fsb_radix_sort_keys_workspace(n, begin_bit, end_bit)
    ws = torch.empty((ws_bytes,), dtype=torch.uint8, device=keys.device)
    in_b = ctypes.c_int(0)
    ev = kernel_timer.start("radix_sort")
    check(lib.fsb_radix_sort_keys(n, ptr(n_dev), begin_bit, end_bit, ptr(keys), ptr(keys_b), ptr(low), ptr(ws), ws_bytes,
                                  ctypes.addressof(in_b), _stream()), "fsb_radix_sort_keys")
    kernel_timer.stop(ev)
    return (keys_b if in_b.value else keys), low


def isect_offsets(sorted_ids: Tensor, C: int, tile_w: int, tile_h: int, n_dev: Optional[Tensor] = None) -> Tensor:
    n_tiles = tile_w * tile_h
    offsets = torch.empty((C, tile_h, tile_w), dtype=torch.int32, device=sorted_ids.device)
    check(lib.fsb_isect_offsets(sorted_ids.numel(), ptr(n_dev), ptr(sorted_ids), C, n_tiles, tile_bits_for(n_tiles),
                                ptr(offsets), _stream()), "fsb_isect_offsets")
    return offsets


def isect_tiles(means2d, radii, depths, tile_size, tile_w, tile_h, tiles_per_gauss=None, legacy_bbox=False,
                sort=True, totals=None, reach=None, lazy_ids=False):
    """gsplat.isect_tiles + isect_offset_encode in one go (unpacked layout).

    means2d [C,N,2], radii [C,N] int32, depths [C,N] -> tiles_per_gauss [C,N], isect_ids [n_isects] int64 (sorted),
    flatten_ids [n_isects] int32, isect_offsets [C,th,tw] int32.
    """
    _req_cuda(means2d, radii, depths)
    C, N = radii.shape
    if tiles_per_gauss is None:
        tiles_per_gauss = isect_count(means2d, radii, tile_size, tile_w, tile_h, legacy_bbox)
    # `reach` = (conics [C,N,3], opacities [C,N]), EXPERIMENTAL: the lists hold only the (Gaussian, tile) pairs that can
    # pass the alpha test somewhere in the tile (csrc/isect_reach.cu); `tiles_per_gauss` stays the bounding-box count
    # the API reports, the lists are built from the reach counts
    list_counts = tiles_per_gauss
    # Two-level binning (pruned lists, sorted): Gaussians by depth first, entries emitted in that order, then a stable
    # sort on the (camera, tile) bits alone — the same order as sorting the 64-bit (camera | tile | depth) keys, with
    # 2 passes over 8-byte entries instead of 5 over 12-byte pairs (include/fsb200.h).
    two_level = reach is not None and sort and TWO_LEVEL_BINNING
    perm = None
    if reach is not None:
        reach = (_f32c(reach[0]), _f32c(reach[1]))
        list_counts = isect_count_reach(means2d, radii, reach[0], reach[1], tile_size, tile_w, tile_h, legacy_bbox,
                                        depths=depths if two_level else None)
        if two_level:
            _, perm = radix_sort_pairs(list_counts.depth_keys, list_counts.depth_vals, 32)
    hit_masks = getattr(list_counts, "hit_masks", None)
    hi_end = sort_end_bit(tile_w * tile_h, C)
    st = static_mode()
    if st is not None:
        # no host read: buffers for the capacity, the true count stays in totals[0] on the device
        if totals is None:
            totals = torch.zeros(1, dtype=torch.int64, device=radii.device)
        offsets, _ = isect_scan(list_counts, totals, read_back=False, perm=perm)
        n_dev = totals[:1]
        st.counts.append(n_dev)
        capacity = st.capacity
    else:
        offsets, n_isects = isect_scan(list_counts, totals, perm=perm)
        if totals is not None:
            totals.host = n_isects  # the values that came back with the one D2H read
            n_isects = n_isects[0]
        n_dev, capacity = None, n_isects
    ids, flat = isect_emit(means2d, radii, depths, offsets, capacity, C, N, tile_size, tile_w, tile_h, legacy_bbox,
                           n_dev=n_dev, overflow=st.overflow if st is not None else None, reach=reach,
                           hit_masks=hit_masks, perm=perm, packed=two_level)
    if two_level:
        ids, flat = radix_sort_keys(ids, 32, hi_end, n_dev=n_dev)
        ids = PackedIsectIds(ids, flat, depths)  # `lazy_ids`: the caller may never ask for the gsplat key form
    elif sort:
        ids, flat = radix_sort_pairs(ids, flat, hi_end, n_dev=n_dev)
    # the (camera, tile) id sits in the high word of either key format
    tile_offsets = isect_offsets(ids.packed if two_level else ids, C, tile_w, tile_h, n_dev=n_dev)
    if n_dev is not None:
        flat.n_dev = n_dev
    if two_level and not lazy_ids:
        ids = ids.materialize()
    return tiles_per_gauss, ids, flat, tile_offsets


TWO_LEVEL_BINNING = os.environ.get("FSB_TWO_LEVEL_BINNING", "1") != "0"


class PackedIsectIds:
    """`isect_ids` of the two-level binning: the sorted keys exist as (camera, tile) << 32 | flatten id; the gsplat form
    (camera, tile) << 32 | depth bits is built on request (`.ids()`, `torch.as_tensor(...)`-style consumers call
    `materialize`).  Only tests and debugging ask for it; the compositing kernels use flatten_ids and isect_offsets."""

    def __init__(self, packed: Tensor, flatten_ids: Tensor, depths: Tensor):
        self.packed, self._flat, self._depths = packed, flatten_ids, depths

    def materialize(self) -> Tensor:
        n_dev = getattr(self._flat, "n_dev", None)
        flat = (self._flat & 0x7FFFFFFF).long()
        if n_dev is not None:  # static-capacity mode: entries past the true count are uninitialised
            flat = flat.clamp_(0, self._depths.numel() - 1)
        depth_bits = self._depths.reshape(-1).view(torch.int32)[flat].long() & 0xFFFFFFFF
        return (self.packed & ~0xFFFFFFFF) | depth_bits

    def numel(self):
        return self.packed.numel()


def isect_tiles_legacy_shared(means2d, radii, depths, tile_size, tile_w, tile_h, first_flat, first_offsets,
                              lists_done=None, reach=None):
    """Static-capacity mode only: binning of the legacy (gsplat 0.1.x bbox rule) normals pass that follows a
    rasterization() on the same projected Gaussians.  The device compares the legacy intersection total with the
    first pass's; when they agree (identical lists) emit / sort / offsets degenerate to no-ops and the first pass's
    sorted lists are copied over (fsb_isect_share_gate / fsb_isect_share_copy), otherwise the legacy lists are
    built as usual.  `lists_done`: event recorded after the first pass's binning (the copy waits for it, so the
    count and scan of this pass may run beside it on another stream).  Returns (flatten_ids, isect_offsets)."""
    st = static_mode()
    assert st is not None and getattr(first_flat, "n_dev", None) is not None
    _req_cuda(means2d, radii, depths)
    C, N = radii.shape
    dev = radii.device
    # `reach`: the first pass's lists were pruned (isect_tiles(reach=...)), so this pass counts the same way: the
    # reached tiles of the legacy box are a superset of the reached tiles of the 1.0 box, equal totals <=> same lists
    if reach is not None:
        reach = (_f32c(reach[0]), _f32c(reach[1]))
        counts = isect_count_reach(means2d, radii, reach[0], reach[1], tile_size, tile_w, tile_h, True)
    else:
        counts = isect_count(means2d, radii, tile_size, tile_w, tile_h, True)
    totals = torch.zeros(1, dtype=torch.int64, device=dev)
    offsets, _ = isect_scan(counts, totals, read_back=False)
    n_list = totals[:1]
    st.counts.append(n_list)
    gate = torch.empty(1, dtype=torch.int64, device=dev)
    check(lib.fsb_isect_share_gate(ptr(first_flat.n_dev), ptr(n_list), ptr(gate), _stream()), "fsb_isect_share_gate")
    ids, flat = isect_emit(means2d, radii, depths, offsets, st.capacity, C, N, tile_size, tile_w, tile_h, True,
                           n_dev=gate, overflow=st.overflow, reach=reach, hit_masks=getattr(counts, "hit_masks", None))
    end_bit = sort_end_bit(tile_w * tile_h, C)
    ids, flat = radix_sort_pairs(ids, flat, end_bit, n_dev=gate)
    tile_offsets = isect_offsets(ids, C, tile_w, tile_h, n_dev=gate)
    if lists_done is not None:
        torch.cuda.current_stream().wait_event(lists_done)
    cap = min(flat.numel(), first_flat.numel())
    check(lib.fsb_isect_share_copy(ptr(gate), ptr(n_list), cap, ptr(first_flat), ptr(first_offsets),
                                   first_offsets.numel(), ptr(flat), ptr(tile_offsets), _stream()),
          "fsb_isect_share_copy")
    flat.n_dev = n_list
    return flat, tile_offsets


def supported_channels(D: int) -> int:
    return lib.fsb_raster_supported_channels(D)


LEGACY_FLAG = 0x80000000  # FSB_LEGACY_FLAG: bit 31 of a flatten id = entry of the 0.1.x list only (union lists)


def raster_dn_fwd(means2d, conics, colors_a, colors_b, opacities, backgrounds_a, backgrounds_b, width, height,
                  tile_size, isect_offsets_t, flatten_ids, ed_channel=-1, n_dev=None):
    """fsb_raster_dn_fwd: colour sets A [C,N,DA] and B [C,N,DB] composited by one walk of the (union) lists.
    -> out_a [C,H,W,DA], out_b [C,H,W,DB], alphas [C,H,W,1], last_ids [C,H,W], workspace."""
    _req_cuda(means2d, conics, colors_a, colors_b, opacities)
    C = isect_offsets_t.shape[0]
    tile_h, tile_w = isect_offsets_t.shape[1], isect_offsets_t.shape[2]
    N = means2d.shape[-2]
    DA, DB = colors_a.shape[-1], colors_b.shape[-1]
    dev = means2d.device
    out_a = torch.empty((C, height, width, DA), dtype=torch.float32, device=dev)
    out_b = torch.empty((C, height, width, DB), dtype=torch.float32, device=dev)
    alphas = torch.empty((C, height, width, 1), dtype=torch.float32, device=dev)
    last_ids = torch.empty((C, height, width), dtype=torch.int32, device=dev)
    ws_bytes = lib.fsb_raster_dn_workspace(flatten_ids.numel(), C * tile_h * tile_w, C * N, DA, DB)
    ws = torch.empty((ws_bytes,), dtype=torch.uint8, device=dev)
    ev = kernel_timer.start(f"raster_fwd_D{DA}+{DB}")
    check(lib.fsb_raster_dn_fwd(C, N, DA, DB, flatten_ids.numel(), ptr(n_dev), ptr(means2d), ptr(conics), ptr(colors_a),
                                ptr(colors_b), ptr(opacities), ptr(backgrounds_a), ptr(backgrounds_b), None, width,
                                height, tile_size, tile_w, tile_h, ptr(isect_offsets_t), ptr(flatten_ids),
                                int(ed_channel), ptr(ws), ws_bytes, ptr(out_a), ptr(out_b), ptr(alphas), ptr(last_ids),
                                _stream()), "fsb_raster_dn_fwd")
    kernel_timer.stop(ev)
    if pair_probe.enabled:
        pair_probe.count(f"D{DA}", C, N, flatten_ids, n_dev, means2d, conics, opacities, width, height, tile_size,
                         tile_w, tile_h, isect_offsets_t, last_ids)
    return out_a, out_b, alphas, last_ids, ws


def raster_dn_bwd(shape_cn, DA, DB, backgrounds_a, backgrounds_b, width, height, tile_size, isect_offsets_t, n_list,
                  ed_channel, ws, render_a, render_alphas, last_ids, v_a, v_b, v_alphas, absgrad, need_xy=True,
                  n_dev=None):
    C = isect_offsets_t.shape[0]
    tile_h, tile_w = isect_offsets_t.shape[1], isect_offsets_t.shape[2]
    Cn, N = shape_cn
    CN = Cn * N
    dev = render_alphas.device
    sizes = [2 * CN if need_xy else 0, 2 * CN if (absgrad and need_xy) else 0, 3 * CN, DA * CN, DB * CN, CN]
    flat = torch.zeros((sum(sizes),), dtype=torch.float32, device=dev)
    parts = torch.split(flat, sizes)
    v_means2d = parts[0].view(Cn, N, 2) if need_xy else None
    v_abs = parts[1].view(Cn, N, 2) if (absgrad and need_xy) else None
    v_conics = parts[2].view(Cn, N, 3)
    v_colors_a = parts[3].view(Cn, N, DA)
    v_colors_b = parts[4].view(Cn, N, DB)
    v_opac = parts[5].view(Cn, N)
    ev = kernel_timer.start(f"raster_bwd_D{DA}+{DB}")
    check(lib.fsb_raster_dn_bwd(C, N, DA, DB, n_list, ptr(n_dev), ptr(backgrounds_a), ptr(backgrounds_b), None, width,
                                height, tile_size, tile_w, tile_h, ptr(isect_offsets_t), int(ed_channel), ptr(ws),
                                ws.numel(), ptr(render_a), ptr(render_alphas), ptr(last_ids), ptr(v_a), ptr(v_b),
                                ptr(v_alphas), ptr(v_abs), ptr(v_means2d), ptr(v_conics), ptr(v_colors_a),
                                ptr(v_colors_b), ptr(v_opac), _stream()), "fsb_raster_dn_bwd")
    kernel_timer.stop(ev)
    return v_means2d, v_abs, v_conics, v_colors_a, v_colors_b, v_opac


def raster_fwd(means2d, conics, colors, opacities, backgrounds, masks, width, height, tile_size, isect_offsets_t,
               flatten_ids, ed_normalize=False, n_dev=None):
    _req_cuda(means2d, conics, colors, opacities)
    C = isect_offsets_t.shape[0]
    tile_h, tile_w = isect_offsets_t.shape[1], isect_offsets_t.shape[2]
    N = means2d.shape[-2]
    D = colors.shape[-1]
    dev = means2d.device
    out = torch.empty((C, height, width, D), dtype=torch.float32, device=dev)
    alphas = torch.empty((C, height, width, 1), dtype=torch.float32, device=dev)
    last_ids = torch.empty((C, height, width), dtype=torch.int32, device=dev)
    ws_bytes = lib.fsb_raster_workspace(flatten_ids.numel(), C * tile_h * tile_w, C * N, D)
    ws = torch.empty((ws_bytes,), dtype=torch.uint8, device=dev)
    ev = kernel_timer.start(f"raster_fwd_D{D}")
    check(lib.fsb_raster_fwd(C, N, D, flatten_ids.numel(), ptr(n_dev), ptr(means2d), ptr(conics), ptr(colors), ptr(opacities),
                             ptr(backgrounds), ptr(masks), width, height, tile_size, tile_w, tile_h,
                             ptr(isect_offsets_t), ptr(flatten_ids), int(ed_normalize), ptr(ws), ws_bytes, ptr(out),
                             ptr(alphas), ptr(last_ids), _stream()), "fsb_raster_fwd")
    kernel_timer.stop(ev)
    if pair_probe.enabled:
        pair_probe.count(f"D{D}", C, N, flatten_ids, n_dev, means2d, conics, opacities, width, height, tile_size,
                         tile_w, tile_h, isect_offsets_t, last_ids)
    return out, alphas, last_ids, ws


class _PairProbe:
    """Measurement aid for bench.py (off by default): after a forward pass, count the (pixel, entry) pairs it blended
    and the pairs a per-pixel list walk visits (fsb_raster_pair_count) — the Q of SURVEY.md §8d's flop figures."""

    def __init__(self):
        self.enabled = False
        self.results = {}

    def count(self, key, C, N, flatten_ids, n_dev, means2d, conics, opacities, width, height, tile_size, tile_w,
              tile_h, isect_offsets_t, last_ids):
        counts = torch.zeros(2, dtype=torch.int64, device=means2d.device)
        check(lib.fsb_raster_pair_count(C, N, flatten_ids.numel(), ptr(n_dev), ptr(means2d), ptr(conics), ptr(opacities),
                                        width, height, tile_size, tile_w, tile_h, ptr(isect_offsets_t),
                                        ptr(flatten_ids), ptr(last_ids), ptr(counts), _stream()),
              "fsb_raster_pair_count")
        self.results[key] = counts

    def summary(self):
        return {k: {"blended": int(v[0]), "visited": int(v[1])} for k, v in self.results.items()}


pair_probe = _PairProbe()


def raster_bwd(means2d, conics, colors, opacities, backgrounds, masks, width, height, tile_size, isect_offsets_t,
               flatten_ids, ed_normalize, ws, render_colors, render_alphas, last_ids, v_render_colors, v_render_alphas,
               absgrad, need_xy=True, n_dev=None):
    C = isect_offsets_t.shape[0]
    tile_h, tile_w = isect_offsets_t.shape[1], isect_offsets_t.shape[2]
    N = means2d.shape[-2]
    D = colors.shape[-1]
    # one zero-filled allocation (one fill launch) carved into the five accumulation targets
    CN = means2d.shape[0] * N
    sizes = [2 * CN if need_xy else 0, 2 * CN if (absgrad and need_xy) else 0, 3 * CN, D * CN, CN]
    flat = torch.zeros((sum(sizes),), dtype=torch.float32, device=means2d.device)
    parts = torch.split(flat, sizes)
    v_means2d = parts[0].view(means2d.shape) if need_xy else None
    v_abs = parts[1].view(means2d.shape) if (absgrad and need_xy) else None
    v_conics = parts[2].view(conics.shape)
    v_colors = parts[3].view(colors.shape)
    v_opac = parts[4].view(opacities.shape)
    ev = kernel_timer.start(f"raster_bwd_D{D}")
    check(lib.fsb_raster_bwd(C, N, D, flatten_ids.numel(), ptr(n_dev), ptr(means2d), ptr(conics), ptr(colors), ptr(opacities),
                             ptr(backgrounds), ptr(masks), width, height, tile_size, tile_w, tile_h,
                             ptr(isect_offsets_t), ptr(flatten_ids), int(ed_normalize), ptr(ws), ws.numel(),
                             ptr(render_colors), ptr(render_alphas), ptr(last_ids), ptr(v_render_colors), ptr(v_render_alphas),
                             ptr(v_abs), ptr(v_means2d), ptr(v_conics), ptr(v_colors), ptr(v_opac), _stream()),
          "fsb_raster_bwd")
    kernel_timer.stop(ev)
    return v_means2d, v_abs, v_conics, v_colors, v_opac


# ---------------------------------------------------------------------------------------------
# autograd Functions
# ---------------------------------------------------------------------------------------------
class ProjectSH(torch.autograd.Function):
    """fully_fused_projection (+ spherical_harmonics + clamp_min(+0.5) + tile count) of gsplat 1.0.0."""

    @staticmethod
    def forward(ctx, means, quats, scales, coeffs, viewmats, Ks, campos, width, height, eps2d, near_plane,
                far_plane, radius_clip, tile_size, sh_degree, color_stride, depth_channel, calc_comp,
                legacy_extra=None):
        means, quats, scales = _f32c(means), _f32c(quats), _f32c(scales)
        viewmats, Ks, coeffs, campos = _f32c(viewmats), _f32c(Ks), _f32c(coeffs), _f32c(campos)
        radii, means2d, depths, conics, comps, colors, tiles = project_sh_fwd(
            means, quats, scales, viewmats, Ks, width, height, eps2d, near_plane, far_plane, radius_clip, tile_size,
            sh_degree, coeffs, campos, color_stride, depth_channel, calc_comp, legacy_extra)
        ctx.save_for_backward(means, quats, scales, coeffs, viewmats, Ks, campos, radii)
        ctx.cfg = (width, height, eps2d, sh_degree, color_stride, depth_channel)
        ctx.set_materialize_grads(False)
        ctx.mark_non_differentiable(radii, tiles)
        return radii, means2d, depths, conics, comps, colors, tiles

    @staticmethod
    def backward(ctx, _v_radii, v_means2d, v_depths, v_conics, v_comps, v_colors, _v_tiles):
        means, quats, scales, coeffs, viewmats, Ks, campos, radii = ctx.saved_tensors
        width, height, eps2d, sh_degree, color_stride, depth_channel = ctx.cfg
        C, N = radii.shape
        dev = means.device
        v_means2d = _f32c(v_means2d) if v_means2d is not None else torch.zeros((C, N, 2), device=dev)
        v_conics = _f32c(v_conics) if v_conics is not None else torch.zeros((C, N, 3), device=dev)
        v_depths, v_comps, v_colors = _f32c(v_depths), _f32c(v_comps), _f32c(v_colors)
        v_means = torch.empty_like(means)
        v_quats = torch.empty_like(quats)
        v_scales = torch.empty_like(scales)
        want_sh = coeffs is not None and v_colors is not None and ctx.needs_input_grad[3]
        v_coeffs = torch.empty_like(coeffs) if want_sh else None
        want_view = ctx.needs_input_grad[4]
        v_viewmats = torch.zeros_like(viewmats) if want_view else None
        want_campos = campos is not None and ctx.needs_input_grad[6] and v_colors is not None
        v_campos = torch.zeros_like(campos) if want_campos else None
        K = coeffs.shape[1] if coeffs is not None else 0
        sh = -1 if (sh_degree is None or v_colors is None) else sh_degree
        if sh >= 0 and v_coeffs is None:
            # the kernel writes SH gradients whenever it evaluates SH; give it a scratch target
            v_coeffs = torch.empty_like(coeffs)
        ev = kernel_timer.start("project_sh_bwd")
        check(lib.fsb_project_sh_bwd(C, N, ptr(means), ptr(quats), ptr(scales), ptr(viewmats), ptr(Ks), width, height,
                                     eps2d, sh, K, ptr(coeffs), ptr(campos), color_stride, depth_channel, ptr(radii),
                                     ptr(v_means2d), ptr(v_depths), ptr(v_conics), ptr(v_comps), ptr(v_colors),
                                     ptr(v_means), ptr(v_quats), ptr(v_scales), ptr(v_coeffs), ptr(v_viewmats),
                                     ptr(v_campos), _stream()), "fsb_project_sh_bwd")
        kernel_timer.stop(ev)
        if coeffs is not None and v_coeffs is None and ctx.needs_input_grad[3]:
            v_coeffs = torch.zeros_like(coeffs)
        return (v_means, v_quats, v_scales, v_coeffs if ctx.needs_input_grad[3] else None, v_viewmats, None, v_campos,
                None, None, None, None, None, None, None, None, None, None, None, None)


class ProjectParams(torch.autograd.Function):
    """ProjectSH fed with the model's parameters as stored (fsb_project_params_fwd / _bwd): un-normalised
    quaternions, log-scales (exp inside) and the SH coefficients as features_dc[N,3] + features_rest[N,K-1,3].
    Saves the activation launches, the [N,K,3] torch.cat of dn_model.py:566-574 and the autograd mirrors of both."""

    @staticmethod
    def forward(ctx, means, quats, log_scales, features_dc, features_rest, viewmats, Ks, width, height, eps2d,
                near_plane, far_plane, radius_clip, tile_size, sh_degree, color_stride, depth_channel,
                legacy_extra=None):
        means, quats, log_scales = _f32c(means), _f32c(quats), _f32c(log_scales)
        features_dc, features_rest = _f32c(features_dc), _f32c(features_rest)
        viewmats, Ks = _f32c(viewmats), _f32c(Ks)
        _req_cuda(means, quats, log_scales, features_dc, features_rest, viewmats, Ks)
        if viewmats.requires_grad:
            raise NotImplementedError("ProjectParams has no view-matrix gradient (camera optimiser off in the "
                                      "reference, dn_model.py:128-130); use rasterization() for that")
        C, N = viewmats.shape[0], means.shape[0]
        assert features_dc.shape == (N, 3), features_dc.shape
        K = 1 + (features_rest.shape[1] if features_rest is not None else 0)
        assert features_rest is None or features_rest.shape == (N, K - 1, 3), features_rest.shape
        dev = means.device
        tile_w, tile_h = math.ceil(width / tile_size), math.ceil(height / tile_size)
        radii = torch.empty((C, N), dtype=torch.int32, device=dev)
        means2d = torch.empty((C, N, 2), dtype=torch.float32, device=dev)
        depths = torch.empty((C, N), dtype=torch.float32, device=dev)
        conics = torch.empty((C, N, 3), dtype=torch.float32, device=dev)
        colors = torch.empty((C, N, color_stride), dtype=torch.float32, device=dev)
        tiles = torch.empty((C, N), dtype=torch.int32, device=dev)
        ev = kernel_timer.start("project_sh_fwd")
        check(lib.fsb_project_params_fwd(C, N, ptr(means), ptr(quats), ptr(log_scales), 1, ptr(viewmats), ptr(Ks),
                                         width, height, eps2d, near_plane, far_plane, radius_clip, tile_size, tile_w,
                                         tile_h, sh_degree, K, ptr(features_dc), ptr(features_rest), color_stride,
                                         depth_channel, ptr(radii), ptr(means2d), ptr(depths), ptr(conics),
                                         ptr(colors), ptr(tiles), ptr(legacy_extra), _stream()),
              "fsb_project_params_fwd")
        kernel_timer.stop(ev)
        ctx.save_for_backward(means, quats, log_scales, features_dc, features_rest, viewmats, Ks, radii)
        ctx.cfg = (width, height, eps2d, sh_degree, K, color_stride, depth_channel)
        ctx.set_materialize_grads(False)
        ctx.mark_non_differentiable(radii, tiles)
        return radii, means2d, depths, conics, colors, tiles

    @staticmethod
    def backward(ctx, _v_radii, v_means2d, v_depths, v_conics, v_colors, _v_tiles):
        means, quats, log_scales, features_dc, features_rest, viewmats, Ks, radii = ctx.saved_tensors
        width, height, eps2d, sh_degree, K, color_stride, depth_channel = ctx.cfg
        C, N = radii.shape
        dev = means.device
        v_means2d = _f32c(v_means2d) if v_means2d is not None else torch.zeros((C, N, 2), device=dev)
        v_conics = _f32c(v_conics) if v_conics is not None else torch.zeros((C, N, 3), device=dev)
        v_colors = _f32c(v_colors) if v_colors is not None else torch.zeros((C, N, color_stride), device=dev)
        v_depths = _f32c(v_depths)
        v_means = torch.empty_like(means)
        v_quats = torch.empty_like(quats)
        v_scales = torch.empty_like(log_scales)
        v_dc = torch.empty_like(features_dc)
        v_rest = torch.empty_like(features_rest) if features_rest is not None else None
        ev = kernel_timer.start("project_sh_bwd")
        check(lib.fsb_project_params_bwd(C, N, ptr(means), ptr(quats), ptr(log_scales), 1, ptr(viewmats), ptr(Ks),
                                         width, height, eps2d, sh_degree, K, ptr(features_dc), ptr(features_rest),
                                         color_stride, depth_channel, ptr(radii), ptr(v_means2d), ptr(v_depths),
                                         ptr(v_conics), ptr(v_colors), ptr(v_means), ptr(v_quats), ptr(v_scales),
                                         ptr(v_dc), ptr(v_rest), _stream()), "fsb_project_params_bwd")
        kernel_timer.stop(ev)
        return (v_means, v_quats, v_scales, v_dc, v_rest) + (None,) * 13


class RasterizeToPixels(torch.autograd.Function):
    """rasterize_to_pixels of gsplat 1.0.0 (optionally with the ED normalisation fused in)."""

    @staticmethod
    def forward(ctx, means2d, conics, colors, opacities, backgrounds, masks, width, height, tile_size,
                isect_offsets_t, flatten_ids, absgrad, ed_normalize, n_dev=None):
        means2d_c, conics_c = _f32c(means2d), _f32c(conics)
        colors_c, opac_c, bg_c = _f32c(colors), _f32c(opacities), _f32c(backgrounds)
        masks_c = masks.contiguous().to(torch.uint8) if masks is not None else None
        out, alphas, last_ids, ws = raster_fwd(means2d_c, conics_c, colors_c, opac_c, bg_c, masks_c, width, height,
                                               tile_size, isect_offsets_t, flatten_ids, ed_normalize, n_dev=n_dev)
        ctx.n_dev = n_dev
        ctx.save_for_backward(means2d, conics_c, colors_c, opac_c, bg_c, masks_c, isect_offsets_t, flatten_ids, out,
                              alphas, last_ids, ws)
        ctx.cfg = (width, height, tile_size, absgrad, ed_normalize)
        ctx.set_materialize_grads(False)
        return out, alphas

    @staticmethod
    def backward(ctx, v_out, v_alphas):
        (means2d, conics, colors, opac, bg, masks, isect_offsets_t, flatten_ids, out, alphas,
         last_ids, ws) = ctx.saved_tensors
        width, height, tile_size, absgrad, ed_normalize = ctx.cfg
        v_out = _f32c(v_out) if v_out is not None else torch.zeros_like(out)
        v_alphas = _f32c(v_alphas) if v_alphas is not None else torch.zeros_like(alphas)
        v_means2d, v_abs, v_conics, v_colors, v_opac = raster_bwd(
            _f32c(means2d), conics, colors, opac, bg, masks, width, height, tile_size, isect_offsets_t, flatten_ids,
            ed_normalize, ws, out, alphas, last_ids, v_out, v_alphas, absgrad, need_xy=ctx.needs_input_grad[0],
            n_dev=ctx.n_dev)
        if absgrad and v_abs is not None:
            # same contract as gsplat: the tensor handed out as meta["means2d"] gets an .absgrad attribute
            means2d.absgrad = v_abs
        v_bg = None
        if bg is not None and ctx.needs_input_grad[4]:
            v_bg = (v_out * (1.0 - alphas)).sum(dim=(1, 2))
        return v_means2d, v_conics, v_colors, v_opac, v_bg, None, None, None, None, None, None, None, None, None


class RasterizeDN(torch.autograd.Function):
    """The two compositing passes of a DN-Splatter iteration in one walk (fsb_raster_dn_fwd / _bwd): colour set A =
    rasterization()'s RGB + depth colours, colour set B = the per-Gaussian normals of the legacy rasterize_gaussians
    pass (dn_model.py:644-653), same means2d / conics / opacities.  The 2-D mean gradient comes from set A only."""

    @staticmethod
    def forward(ctx, means2d, conics, colors_a, colors_b, opacities, backgrounds_a, backgrounds_b, width, height,
                tile_size, isect_offsets_t, flatten_ids, absgrad, ed_channel, n_dev=None):
        means2d_c, conics_c = _f32c(means2d), _f32c(conics)
        ca, cb, opac_c = _f32c(colors_a), _f32c(colors_b), _f32c(opacities)
        bga, bgb = _f32c(backgrounds_a), _f32c(backgrounds_b)
        out_a, out_b, alphas, last_ids, ws = raster_dn_fwd(means2d_c, conics_c, ca, cb, opac_c, bga, bgb, width,
                                                           height, tile_size, isect_offsets_t, flatten_ids, ed_channel,
                                                           n_dev=n_dev)
        ctx.n_dev = n_dev
        ctx.save_for_backward(means2d, bga, bgb, isect_offsets_t, out_a, alphas, last_ids, ws)
        ctx.cfg = (width, height, tile_size, absgrad, ed_channel, tuple(opac_c.shape), ca.shape[-1], cb.shape[-1],
                   flatten_ids.numel())
        ctx.set_materialize_grads(False)
        return out_a, out_b, alphas

    @staticmethod
    def backward(ctx, v_a, v_b, v_alphas):
        means2d, bga, bgb, isect_offsets_t, out_a, alphas, last_ids, ws = ctx.saved_tensors
        width, height, tile_size, absgrad, ed_channel, shape_cn, DA, DB, n_list = ctx.cfg
        v_a = _f32c(v_a) if v_a is not None else torch.zeros_like(out_a)
        v_b = _f32c(v_b) if v_b is not None else out_a.new_zeros(out_a.shape[:-1] + (DB,))
        v_alphas = _f32c(v_alphas) if v_alphas is not None else torch.zeros_like(alphas)
        v_means2d, v_abs, v_conics, v_ca, v_cb, v_opac = raster_dn_bwd(
            shape_cn, DA, DB, bga, bgb, width, height, tile_size, isect_offsets_t, n_list, ed_channel, ws, out_a,
            alphas, last_ids, v_a, v_b, v_alphas, absgrad, need_xy=ctx.needs_input_grad[0], n_dev=ctx.n_dev)
        if absgrad and v_abs is not None:
            means2d.absgrad = v_abs  # same contract as gsplat: meta["means2d"] gets an .absgrad attribute
        return (v_means2d, v_conics, v_ca, v_cb, v_opac) + (None,) * 10
