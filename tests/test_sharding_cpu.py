"""Batch sharding: shard arithmetic, global-batch RNG replay, and the optional output gather over a
2-rank gloo group on CPU.  The data path has no collective; only the gather communicates."""
import os
import socket

import pytest
import torch
import torch.multiprocessing as mp
import torch.nn as nn

from cognitive_aim_depth_estimation_b200 import sharding
from cognitive_aim_depth_estimation_b200.model import create_model

CFG = {"model": {"cognitive_modules": ["ambient_stream", "iterative_focal_stream", "exif_prior_database"]}}


def test_shard_bounds_cover_batch_contiguously():
    for n in (0, 1, 7, 32, 64, 511, 512):
        for world in (1, 2, 3, 4, 8):
            b = sharding.shard_bounds(n, world)
            assert len(b) == world and b[0][0] == 0 and b[-1][1] == n
            assert all(b[i][1] == b[i + 1][0] for i in range(world - 1))
            sizes = [hi - lo for lo, hi in b]
            assert max(sizes) - min(sizes) <= 1 and sizes == sorted(sizes, reverse=True)
    assert sharding.shard_bounds(512, 8) == [(64 * r, 64 * r + 64) for r in range(8)]  # BASELINE config 4
    with pytest.raises(ValueError):
        sharding.shard_range(8, 2, 2)
    with pytest.raises(ValueError):
        sharding.shard_bounds(8, 0)


def test_shard_batch_slices_images_and_exif():
    x = torch.arange(10 * 3 * 2 * 2, dtype=torch.float32).reshape(10, 3, 2, 2)
    ex = {"focal_length": torch.arange(10.0), "camera_idx": torch.arange(10).reshape(10, 1)}
    parts = [sharding.shard_batch(x, ex, r, 3) for r in range(3)]
    assert torch.equal(torch.cat([p[0] for p in parts]), x)
    assert torch.equal(torch.cat([p[1]["camera_idx"] for p in parts]), ex["camera_idx"])
    assert [p[0].shape[0] for p in parts] == [4, 3, 3]
    assert sharding.shard_batch(x, None, 0, 2)[1] is None


def test_rng_replay_uses_global_batch():
    """A shard of 4 images out of a global batch of 8 must draw the per-call projection (src/model.py:1421) that the
    un-sharded call of 8 images draws: randn(8,192), randn(8,768), then nn.Linear(768,64)."""
    m = create_model(CFG, {"num_cameras": 71})
    torch.manual_seed(11)
    torch.randn(8, 192)
    torch.randn(8, 768)
    want = nn.Linear(768, 64)
    m.rng_replay_batch, m.rng_replay_offset = 8, 4
    torch.manual_seed(11)
    eps, noise = m._curiosity_draw(4)
    got = nn.Linear(768, 64)
    assert torch.equal(got.weight, want.weight) and torch.equal(got.bias, want.bias)
    # ... and the shard's CuriosityModule sees ITS rows of the global draws (src/model.py:609,744)
    torch.manual_seed(11)
    assert torch.equal(eps, torch.randn(8, 192)[4:8]) and torch.equal(noise, torch.randn(8, 768)[4:8])
    m.rng_replay_batch, m.rng_replay_offset = None, 0
    torch.manual_seed(11)
    m._curiosity_draw(4)
    other = nn.Linear(768, 64)
    assert not torch.equal(other.weight, want.weight)


class _HostOnlyModel:
    """Stands in for the CUDA model in the gloo test: same host protocol (rng_replay_batch, per-call projection from
    the CPU generator, per-image outputs), trivial per-image arithmetic."""
    rng_replay_batch = None
    rng_replay_offset = 0

    def forward_with_guidance(self, images, exif_data=None, attention_guidance=None, return_attention=False):
        B = self.rng_replay_batch or images.shape[0]
        torch.randn(B, 192)
        torch.randn(B, 768)
        proj = nn.Linear(768, 64)
        feat = images.reshape(images.shape[0], -1)[:, :768] @ proj.weight.detach().t()
        depth = feat.sum(1, keepdim=True) + exif_data["focal_length"].reshape(-1, 1)
        conf = torch.sigmoid(feat.mean(1, keepdim=True))
        heat = torch.softmax(feat[:, :16], dim=-1)
        return (depth, conf, heat) if return_attention else (depth, conf)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, n, out_dir):
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        g = torch.Generator().manual_seed(5)
        x = torch.randn(n, 3, 16, 16, generator=g)
        ex = {"focal_length": torch.rand(n, generator=g) * 100}
        model = _HostOnlyModel()
        torch.manual_seed(11)
        want = model.forward_with_guidance(x, ex, "center", return_attention=True)  # un-sharded, global batch
        runner = sharding.ShardedInference(model, rank, world, gather=True)
        torch.manual_seed(11)
        got = runner.forward_with_guidance(x, ex, "center", return_attention=True)
        assert model.rng_replay_batch is None and model.rng_replay_offset == 0
        ok = all(torch.equal(a, b) for a, b in zip(got, want)) and got[2].shape == (n, 16)
        # local-only mode returns just this rank's rows
        lo, hi = sharding.shard_range(n, rank, world)
        torch.manual_seed(11)
        loc = sharding.ShardedInference(model, rank, world).forward_with_guidance(x, ex, "center")
        ok = ok and torch.equal(loc[0], want[0][lo:hi]) and len(loc) == 2
        open(os.path.join(out_dir, f"rank{rank}.ok" if ok else f"rank{rank}.bad"), "w").close()
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("n", [8, 7])  # even and ragged shards
def test_sharded_equals_unsharded_over_gloo(tmp_path, n):
    world = 2
    mp.spawn(_worker, args=(world, _free_port(), n, str(tmp_path)), nprocs=world, join=True)
    assert sorted(os.listdir(tmp_path)) == ["rank0.ok", "rank1.ok"]
