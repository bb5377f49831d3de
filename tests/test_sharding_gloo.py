"""World-size-2 gloo test of the host-side batch sharding (no GPU): shards cover the batch exactly once, ragged
batches work, and the pose all-gather restores batch order."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from handmvnet_b200.sharding import gather_poses, shard_bounds, shard_inputs
from handmvnet_b200.checkpoint import remap_legacy_keys


def test_shard_bounds_cover_batch():
    for batch in (0, 1, 5, 64, 65, 4096):
        for world in (1, 2, 4, 8):
            spans = [shard_bounds(batch, world, r) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == batch
            assert all(spans[i][1] == spans[i + 1][0] for i in range(world - 1))
            assert max(h - l for l, h in spans) - min(h - l for l, h in spans) <= 1


def _worker(rank, world, port, batch):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    x = torch.arange(batch, dtype=torch.float32).reshape(batch, 1, 1, 1, 1).expand(batch, 5, 3, 2, 2)
    xs, _, _ = shard_inputs(x, None, None, world, rank)
    local = xs[:, 0, 0, 0, 0].reshape(-1, 1, 1).expand(-1, 21, 3).contiguous()     # "pose" = sample id
    full = gather_poses(local, batch)
    assert full.shape == (batch, 21, 3)
    assert torch.equal(full[:, 0, 0], torch.arange(batch, dtype=torch.float32))
    dist.destroy_process_group()


def test_gather_poses_world2_gloo():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    mp.spawn(_worker, args=(2, port, 7), nprocs=2, join=True)


def test_legacy_checkpoint_key_remap():
    sd = {"pose_net.conv.0.weight": 1, "sample_net.conv.0.weight": 2, "backbone.conv1.weight": 3}
    out = remap_legacy_keys(sd)
    assert set(out) == {"pose_net.0.weight", "sample_nets.0.conv.0.weight", "backbone.conv1.weight"}
