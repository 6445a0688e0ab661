"""Stand-alone `forward` of every module on the reference's call surface (set_transformer.py:47-216, mlp.py,
activation.py, normalization.py) against goldens of the unmodified reference modules (oracle/make_golden.py --modules),
with randomised AdaGN weights and alphas.  Each forward runs the C-ABI kernels through
gecco_b200/models/_native.py: bf16 tensor-core operands, fp32 accumulation -> rms tolerance 1.5e-2 (exact ops: 1e-5)."""
from pathlib import Path

import pytest
import torch

from tests import synth

pytestmark = pytest.mark.gpu
GOLD = Path(__file__).parent / "golden" / "modules.pt"


def rel(a, b):
    return ((a.double().cpu() - b.double()).pow(2).mean().sqrt() / b.double().pow(2).mean().sqrt()).item()


def sub(t):
    return t[:, ::5, ::8]


def test_module_forwards(cuda):
    from gecco_b200.models import MLP, GaussianActivation, SetTransformer

    g = torch.load(GOLD, weights_only=False)
    r = g["recipe"]
    L = r["n_layers"]
    st = SetTransformer(n_layers=L, num_inducers=synth.NUM_INDUCERS, feature_dim=synth.FEATURE_DIM, t_embed_dim=1,
                        num_heads=synth.NUM_HEADS, activation=GaussianActivation)
    pre = "backbone.model.inner."
    sd = {k[len(pre):]: v for k, v in synth.synth_state_dict(synth.network_shapes("uncond", n_layers=L), r["weight_seed"]).items()
          if k.startswith(pre)}
    st.load_state_dict(sd)
    st = st.to(cuda).eval()
    B, N, C = r["B"], r["N"], synth.FEATURE_DIM
    x = torch.randn(B, N, C, generator=synth.gen(r["x_seed"])).to(cuda)
    t = (torch.randn(B, 1, 1, generator=synth.gen(r["t_seed"])) * r["t_scale"]).to(cuda)
    x2 = torch.randn(B, r["N2"], C, generator=synth.gen(r["x2_seed"])).to(cuda)
    lay = st.layers[0]
    errs = {}
    errs["act"] = rel(sub(lay.mlp[1](x)), g["act"])
    errs["act_raw"] = rel(sub(GaussianActivation(normalized=False).to(cuda)(x)), g["act_raw"])
    errs["adagn"] = rel(sub(lay.broadcast_norm(x, t)), g["adagn"])
    assert errs["act"] < 1e-5 and errs["act_raw"] < 1e-5 and errs["adagn"] < 1e-5, errs
    errs["mlp"] = rel(sub(lay.mlp(x)), g["mlp"])
    rm = r["relu_mlp"]
    relu_mlp = MLP(C, rm["out"], rm["width"], depth=rm["depth"])
    relu_mlp.load_state_dict(synth.synth_state_dict({k: tuple(v.shape) for k, v in relu_mlp.state_dict().items()}, rm["seed"]))
    relu_mlp = relu_mlp.to(cuda).eval()
    errs["relu_mlp"] = rel(sub(relu_mlp(x)), g["relu_mlp"])
    assert relu_mlp(x[0, :7]).shape == (7, rm["out"])  # arbitrary leading dimensions like nn.Sequential
    errs["pool"] = rel(sub(lay.broadcast.pool(x)), g["pool"])
    attn, h = lay.broadcast(x, t, return_h=True)
    errs["broadcast"], errs["broadcast_h"] = rel(sub(attn), g["broadcast"]), rel(sub(h), g["broadcast_h"])
    attn2, none = lay.broadcast(x2, t, return_h=False, h=h)
    assert none is None and attn2.shape == x2.shape
    errs["broadcast_cached"] = rel(sub(attn2), g["broadcast_cached"])
    y, h1 = lay(x, t, return_h=True)
    errs["layer"], errs["layer_h"] = rel(sub(y), g["layer"]), rel(sub(h1), g["layer_h"])
    f, hs = st(x, t, return_h=True)
    errs["st"] = rel(sub(f), g["st"])
    assert len(hs) == L
    for l, (a, b) in enumerate(zip(hs, g["st_hs"])):
        errs[f"st_h{l}"] = rel(sub(a), b)
    f2, none = st(x2, t, return_h=False, hs=hs)
    assert none is None
    errs["st_cached"] = rel(sub(f2), g["st_cached"])
    f3, none = st(x, t)
    assert none is None and torch.equal(f3, f)
    print({k: f"{v:.2e}" for k, v in errs.items()})
    bad = {k: v for k, v in errs.items() if v >= 1.5e-2}
    assert not bad, bad
