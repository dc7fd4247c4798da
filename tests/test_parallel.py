"""CPU: the multi-GPU plumbing (shard plan, the one weight broadcast, result gather) on gloo, world_size 2."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from transformer_explainability_b200 import parallel


def test_shard_range_partitions():
    for total in (0, 1, 7, 64, 256, 257):
        for world in (1, 2, 3, 8):
            spans = [parallel.shard_range(total, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == total
            for (a, b), (c, d) in zip(spans, spans[1:]):
                assert b == c
            sizes = [b - a for a, b in spans]
            assert max(sizes) - min(sizes) <= 1


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, total):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world),
                      LOCAL_RANK=str(rank))
    r, w, _ = parallel.init_from_env(backend="gloo")
    assert (r, w) == (rank, world)
    weights = torch.arange(1000, dtype=torch.float32) if rank == 0 else torch.zeros(1000)
    parallel.broadcast_flat_weights(weights, src=0)
    assert torch.equal(weights, torch.arange(1000, dtype=torch.float32))

    class FakeEngine:                      # stands in for ViTEngine: "explains" by a per-sample function
        def explain(self, images, index=None, start_layer=0, chunk=None):
            return images.reshape(images.shape[0], -1)[:, :4] * 2 + weights[:4], torch.zeros(images.shape[0])

    images = torch.arange(total * 6, dtype=torch.float32).reshape(total, 6)
    maps, _ = parallel.explain_sharded(FakeEngine(), images, gather=True)
    expect = images[:, :4] * 2 + weights[:4]
    assert torch.equal(maps, expect)
    dist.barrier()
    dist.destroy_process_group()


def test_gloo_world2_shard_broadcast_gather():
    port = _free_port()
    mp.spawn(_worker, args=(2, port, 7), nprocs=2, join=True)
