"""Host-side logic of the training step on CPU: flat parameter / gradient buffers and the bucketed, backward-overlapped
gradient all-reduce over two gloo ranks (the N > 1 path of BASELINE config 5; NCCL on the GPUs)."""
import copy
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp
from torch import nn


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _toy():
    torch.manual_seed(3)
    # the last Linear is never used in forward: its bucket has to be reduced by finish()
    return nn.ModuleDict(dict(a=nn.Linear(7, 33), b=nn.Linear(33, 65), c=nn.Linear(65, 5), unused=nn.Linear(4, 4)))


def _loss(m, x):
    return m["c"](torch.tanh(m["b"](torch.tanh(m["a"](x))))).pow(2).mean()


def _data(rank):
    return torch.randn(6, 7, generator=torch.Generator().manual_seed(100 + rank))


def test_flat_state_views_accumulate_in_place():
    from gecco_b200.training import FlatState

    m = _toy()
    ref = copy.deepcopy(m)
    st = FlatState(m.parameters())
    assert st.numel % 64 == 0 and all(o % 64 == 0 for o in st.offsets)
    for p, q in zip(m.parameters(), ref.parameters()):
        assert torch.equal(p, q) and p.data_ptr() >= st.p.data_ptr() and p.grad.data_ptr() >= st.g.data_ptr()
    _loss(m, _data(0)).backward()
    _loss(ref, _data(0)).backward()
    for i, (p, q) in enumerate(zip(m.parameters(), ref.parameters())):
        o = st.offsets[i]
        if q.grad is None:
            assert st.g[o:o + p.numel()].abs().max() == 0
        else:
            assert torch.equal(st.g[o:o + p.numel()].view(p.shape), q.grad)  # accumulated into the flat buffer
    m.zero_grad(set_to_none=True)
    st.zero_grad()
    assert st.g.abs().max() == 0 and all(p.grad is not None for p in m.parameters())
    # writes through the flat parameter buffer are the module's weights
    st.p.mul_(2.0)
    for p, q in zip(m.parameters(), ref.parameters()):
        assert torch.equal(p, 2 * q)


def _worker(rank, world, port, results):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from gecco_b200.training import FlatState, GradReducer

        m = _toy()
        refs = [copy.deepcopy(m) for _ in range(world)]
        st = FlatState(m.parameters())
        red = GradReducer(st.params, st.offsets, st.g, bucket_bytes=4 * 600)  # several buckets
        assert len(red.buckets) >= 3 and red.buckets[-1][0] == 0 and red.buckets[0][1] == st.numel
        assert sorted(red.buckets) == sorted(red.buckets, key=lambda b: b[0]) and sum(hi - lo for lo, hi in red.buckets) == st.numel
        for step in range(2):  # the reducer re-arms itself
            st.zero_grad()
            _loss(m, _data(rank) + step).backward()
            scale = red.finish()
            assert scale == 1.0 / world
            for r in range(world):
                refs[r].zero_grad(set_to_none=True)
                _loss(refs[r], _data(r) + step).backward()
            for i, p in enumerate(st.params):
                gs = [list(refs[r].parameters())[i].grad for r in range(world)]
                want = torch.zeros_like(p) if gs[0] is None else sum(gs) / world
                got = st.g[st.offsets[i]:st.offsets[i] + p.numel()].view(p.shape) * scale
                assert torch.allclose(got, want, rtol=1e-6, atol=1e-7), (step, i)
        red.remove()
        results[rank] = True
    finally:
        dist.destroy_process_group()


def test_bucketed_gradient_allreduce_gloo():
    world = 2
    mgr = mp.Manager()
    results = mgr.dict()
    mp.spawn(_worker, args=(world, _free_port(), results), nprocs=world, join=True)
    assert all(results.get(r) for r in range(world))
