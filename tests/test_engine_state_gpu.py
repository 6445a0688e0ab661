"""Host-side state of the engine: the packed feature pyramid and the packed weights must follow what the
caller passes NOW, not what happened to live at the same address before; the sampler's CUDA graph must reproduce the
eager launch sequence bit for bit; CUDA-generator noise is drawn exactly like the reference draws it."""
import pytest
import torch

from tests import synth
from tests.models_b200 import FreshConditioner, build

pytestmark = pytest.mark.gpu


def _cond_model(cuda, feats, n_layers=2):
    rp = synth.SHAPENET_VOL_REPARAM
    m = build("cond", "gaussian", rp["mean"], rp["sigma"], 165.0, 1234, cuda, feats, n_layers=n_layers)
    return m


def _ctx(cuda, B):
    import gecco_b200 as G

    return G.Context3d(image=torch.zeros(B, 3, 8, 8, device=cuda), K=synth.camera(B, synth.K_SHAPENET).to(cuda))


def test_new_image_same_shape_is_not_served_from_the_cache(cuda):
    B, N = 2, 200
    fa = [f.to(cuda) for f in synth.synth_features(B, (34, 17, 8), 1)]
    fb = [f.to(cuda) for f in synth.synth_features(B, (34, 17, 8), 2)]
    ctx = _ctx(cuda, B)
    x = (torch.randn(B, N, 3, generator=synth.gen(3)) * 2).to(cuda)
    sig = torch.tensor([0.5, 5.0], device=cuda)
    model = _cond_model(cuda, None)
    outs = {}
    for tag, feats in (("a", fa), ("b", fb), ("a2", fa)):
        model.conditioner = FreshConditioner(feats)  # new tensors on every call, freed on return: addresses get recycled
        outs[tag] = model(x, sig, ctx).clone()
        outs[tag + "_s"] = model.sample_stochastic((B, N, 3), ctx, rng=synth.gen(4), num_steps=2).clone()
    ref_b = _cond_model(cuda, fb)  # a fresh model / engine that only ever saw image b
    assert torch.equal(outs["b"], ref_b(x, sig, ctx))
    assert torch.equal(outs["b_s"], ref_b.sample_stochastic((B, N, 3), ctx, rng=synth.gen(4), num_steps=2))
    assert not torch.equal(outs["a"], outs["b"]) and torch.equal(outs["a"], outs["a2"]) and torch.equal(outs["a_s"], outs["a2_s"])
    # the batch / shape checks run on every call, cache hit or not
    with pytest.raises(ValueError):
        model(torch.cat([x, x]), torch.cat([sig, sig]), _ctx(cuda, 2 * B))


def test_in_place_weight_update_through_data_is_seen(cuda):
    """`p.data.copy_()` (the reference's EMA swap, ema.py:327-337) neither moves the storage nor bumps `_version`."""
    B, N = 2, 256
    rp = synth.UNCOND_REPARAM
    model = build("uncond", "gaussian", rp["mean"], rp["sigma"], 165.0, 1234, cuda, n_layers=2)
    x = (torch.randn(B, N, 3, generator=synth.gen(5)) * 3).to(cuda)
    sig = torch.tensor([0.3, 9.0], device=cuda)
    d0 = model(x, sig, None).clone()
    sd = {k: v.clone() for k, v in model.state_dict().items()}
    with torch.no_grad():
        for name, p in model.named_parameters():
            if name.endswith("mlp.2.weight") or name.endswith("kv_proj.weight") or name.endswith("mlp_norm.scale.bias"):
                v0 = p._version
                p.data.copy_(p.data * 1.25)
                assert p._version == v0  # the hole the content fingerprint closes
                sd[name] = p.detach().clone()
    d1 = model(x, sig, None)
    fresh = build("uncond", "gaussian", rp["mean"], rp["sigma"], 165.0, None, cuda, n_layers=2, state_dict=sd)
    assert torch.equal(d1, fresh(x, sig, None)) and not torch.equal(d0, d1)


def test_sampler_graph_matches_eager_bitwise(cuda):
    """gecco_sample captured into a CUDA graph (include/gecco_b200.h promises capturability) == the eager launch sequence."""
    from gecco_b200 import _abi
    from gecco_b200.engine import engine_for

    B, N = 3, 640
    feats = synth.synth_features(B, (34, 17, 8), 7)
    model = _cond_model(cuda, feats, n_layers=3)
    ctx = _ctx(cuda, B)
    lib = _abi.load()
    eng = engine_for(*model._network())
    try:
        _abi.check(lib.gecco_set_option(b"graphs", 0))
        eager = model.sample_stochastic((B, N, 3), ctx, rng=synth.gen(8), num_steps=4)
        assert eng.graph_status() == 0
        _abi.check(lib.gecco_set_option(b"graphs", 1))
        cap = model.sample_stochastic((B, N, 3), ctx, rng=synth.gen(8), num_steps=4)
        assert eng.graph_status() == 1, eng.graph_status()
        rep = model.sample_stochastic((B, N, 3), ctx, rng=synth.gen(8), num_steps=4)
        assert eng.graph_status() == 2
        # on a side stream too (the caller's stream is forked into / joined from the engine's capture stream)
        st = torch.cuda.Stream()
        with torch.cuda.stream(st):
            side = model.sample_stochastic((B, N, 3), ctx, rng=synth.gen(8), num_steps=4)
        st.synchronize()
        # a different schedule is a different graph, not a stale replay
        other = model.sample_stochastic((B, N, 3), ctx, rng=synth.gen(8), num_steps=4, sigma_max=80.0)
        assert eng.graph_status() == 1
    finally:
        lib.gecco_set_option(b"graphs", 1)
    assert torch.equal(eager, cap) and torch.equal(eager, rep) and torch.equal(eager, side)
    assert not torch.equal(eager, other)


def test_cuda_generator_noise_matches_reference_draw_order(cuda):
    """The reference draws `torch.randn(shape, generator=rng)` once per step (diffusion.py:306,324); the drop-in fills
    slices of one buffer with `normal_(generator=rng)`.  Same Philox stream, same values."""
    shape = (3, 333, 3)
    g1 = torch.Generator(cuda).manual_seed(42)
    ref = [torch.randn(shape, device=cuda, generator=g1) for _ in range(5)]
    g2 = torch.Generator(cuda).manual_seed(42)
    lat = torch.randn(shape, device=cuda, generator=g2)
    buf = torch.empty((4, *shape), device=cuda)
    for i in range(4):
        buf[i].normal_(generator=g2)
    assert torch.equal(lat, ref[0])
    for i in range(4):
        assert torch.equal(buf[i], ref[i + 1])
