"""ResidentViewFeeder against the contract of DNSplatterDataManager.next_train
(/root/reference/dn_splatter/dn_datamanager.py:96-148): index order and refill, shapes, camera metadata, rank sharding."""
import torch


class _Cams:
    """Minimal stand-in for nerfstudio Cameras slicing."""

    def __init__(self, ids, metadata=None):
        self.ids, self.metadata = ids, metadata

    def __getitem__(self, sl):
        return _Cams(self.ids[sl], None)

    def to(self, device):
        return self


def _cache(n=5, H=12, W=16):
    g = torch.Generator().manual_seed(0)
    out = []
    for i in range(n):
        out.append({"image": torch.rand(H, W, 3, generator=g), "mask": (torch.rand(H, W, generator=g) > 0.5),
                    "sensor_depth": torch.rand(H // 2, W // 2, 1, generator=g), "normal": torch.rand(H, W, 3, generator=g),
                    "image_idx": i})
    return out


def test_next_train_order_refill_and_shapes():
    from fusionsense_b200.datamanager import ResidentViewFeeder

    cache = _cache()
    f = ResidentViewFeeder(cache, _Cams(list(range(5))), device="cpu")
    seen = []
    for step in range(12):
        cam, batch = f.next_train(step)
        seen.append(cam.metadata["cam_idx"])
        assert cam.ids == [seen[-1]] and f.image_idx == seen[-1]
        assert batch["image"].shape == (12, 16, 3) and batch["mask"].shape == (12, 16, 1)
        assert batch["sensor_depth"].shape == (12, 16, 1)  # resized to the image (dn_datamanager.py:111-117)
        assert batch["normal"].shape == (12, 16, 3)
    assert seen == [0, 1, 2, 3, 4, 0, 1, 2, 3, 4, 0, 1]  # sequential pop, refilled when empty (:100-102)
    # the resize is TF.resize's bilinear (no antialias)
    ref = torch.nn.functional.interpolate(cache[2]["sensor_depth"].permute(2, 0, 1)[None], size=(12, 16), mode="bilinear",
                                          align_corners=False)[0].permute(1, 2, 0)
    assert torch.equal(f.batches[2]["sensor_depth"], ref)
    # the batch dict is fresh, the tensors are the resident ones
    _, b1 = f.next_train(12)
    b1["normal"] = None
    assert f.batches[2]["normal"] is not None and f.targets()[2]["image"] is f.batches[2]["image"]


def test_rank_sharding_partitions_the_sequence():
    from fusionsense_b200.datamanager import ResidentViewFeeder

    cache = _cache(n=9)
    world = 4
    feeders = [ResidentViewFeeder(cache, _Cams(list(range(9))), device="cpu", world_size=world, rank=r) for r in range(world)]
    per_step = []
    for step in range(5):
        per_step.append([f.next_train(step)[0].metadata["cam_idx"] for f in feeders])
    flat = [i for row in per_step for i in row]
    assert flat == [i % 9 for i in range(20)]  # the ranks together walk the reference's single sequence
    assert all(len(set(row)) == world for row in per_step)  # no view twice inside one iteration
