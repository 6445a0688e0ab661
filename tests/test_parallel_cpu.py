"""world_size-2 gloo test of the multi-rank sampling plumbing (batch sharding, final gather, max-over-ranks timing)."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, results):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from gecco_b200 import parallel as P
        from gecco_b200.structs import Context3d

        assert P.world() == (rank, world)
        B, N = 6, 5
        ctx = Context3d(image=torch.arange(B * 3.0).reshape(B, 3, 1, 1), K=torch.arange(B * 9.0).reshape(B, 3, 3))

        def fake_sampler(shape, c, seed):
            # cloud b is filled with its global index, recovered from the context slice; the seed must be rank specific
            assert seed == 100 + rank and c.image.shape[0] == shape[0]
            idx = c.K[:, 0, 0] / 9.0
            return idx[:, None, None].expand(shape).clone().double()

        out = P.sample_sharded(fake_sampler, (B, N, 3), ctx, seed=100)
        assert out.shape == (B, N, 3)
        assert torch.equal(out[:, 0, 0], torch.arange(B, dtype=torch.float64))
        assert P.max_over_ranks(10.0 + rank) == 10.0 + world - 1
        with pytest.raises(ValueError):
            P.shard_range(7, rank, world)
        results[rank] = True
    finally:
        dist.destroy_process_group()


def test_sharded_sampling_gloo():
    world = 2
    mgr = mp.Manager()
    results = mgr.dict()
    mp.spawn(_worker, args=(world, _free_port(), results), nprocs=world, join=True)
    assert all(results.get(r) for r in range(world))


def test_single_process_passthrough():
    from gecco_b200 import parallel as P

    assert P.world() == (0, 1)
    assert P.shard_range(8, 0, 1) == (0, 8)
    t = torch.ones(2, 3, 3)
    assert P.gather_clouds(t) is t
    assert P.max_over_ranks(1.5) == 1.5
    assert P.shard_context(None, 0, 1) is None
