"""FC inference kernels + support codec (through the C ABI) vs reference outputs and the oracle."""
import numpy as np
import pytest
import torch

import _tables as T
from _configs import product_config
from oracle import networks as onet

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def _weights(z, tag):
    pre = tag + "/w/"
    return {k[len(pre):]: torch.tensor(z[k]) for k in z.files if k.startswith(pre)}


def _model(tag):
    from muzero_hypermodel_b200 import models
    cfg = product_config({"cartpole_shipped": "cartpole"}.get(tag, tag))
    net = models.MuZeroNetwork(cfg)
    z = T.load("net")
    sd = _weights(z, tag)
    assert list(net.state_dict().keys()) == list(sd.keys())          # reference checkpoint keys, same order
    net.set_weights(sd)
    return net.to(DEV).eval(), cfg, z


@pytest.mark.parametrize("tag", ["cartpole_shipped", "cartpole", "tictactoe_fc"])
def test_fc_matches_reference_outputs(tag):
    from muzero_hypermodel_b200 import models
    net, cfg, z = _model(tag)
    obs = torch.tensor(z[tag + "/obs"], device=DEV)
    v0, r0, p0, s0 = net.initial_inference(obs)
    v1, r1, p1, s1 = net.recurrent_inference(s0, torch.tensor(z[tag + "/act"], device=DEV))
    v2, r2, p2, s2 = net.recurrent_inference(s1, torch.tensor(z[tag + "/act2"], device=DEV))
    for got, name in ((v0, "v0"), (p0, "p0"), (s0, "s0"), (v1, "v1"), (r1, "r1"), (p1, "p1"), (s1, "s1"),
                      (v2, "v2"), (r2, "r2"), (p2, "p2"), (s2, "s2")):
        np.testing.assert_allclose(got.cpu().numpy(), z[f"{tag}/{name}"], rtol=1e-5, atol=2e-6, err_msg=name)
    np.testing.assert_array_equal(r0.cpu().numpy(), z[tag + "/r0"])          # log one-hot: -inf / 0 exactly
    # scalars: ill-conditioned inverse transform, see DESIGN.md "support codec conditioning"
    sv1 = models.support_to_scalar(v1, cfg.support_size)
    np.testing.assert_allclose(sv1.cpu().numpy(), z[tag + "/sv1"], rtol=3e-4, atol=2e-4)
    sr0 = models.support_to_scalar(r0, cfg.support_size)
    assert float(sr0.abs().max()) == 0.0


@pytest.mark.parametrize("tag", ["cartpole", "tictactoe_fc"])
def test_fc_fused_outputs_vs_oracle_large_batch(tag):
    """Scalars / priors straight from the kernel epilogue vs the numpy oracle at B=4096 (+ragged tail)."""
    net, cfg, z = _model(tag)
    A = len(cfg.action_space)
    onn = onet.FullyConnected({k: v.numpy() for k, v in net.get_weights().items()}, A, cfg.support_size)
    rs = np.random.RandomState(7)
    B = 4096 + 37
    obs = rs.uniform(-1, 1, (B,) + tuple(cfg.observation_shape)).astype(np.float32)
    legal = rs.uniform(size=(B, A)) < 0.7
    legal[:, 0] = True
    out = net.initial_inference_fused(torch.tensor(obs, device=DEV), legal=torch.tensor(legal, device=DEV), want_logits=True)
    ov, orw, op, os_ = onn.initial_inference(obs)
    np.testing.assert_allclose(out["state"].cpu().numpy(), os_, rtol=1e-5, atol=2e-6)
    np.testing.assert_allclose(out["policy_logits"].cpu().numpy(), op, rtol=1e-5, atol=2e-6)
    np.testing.assert_allclose(out["value_logits"].cpu().numpy(), ov, rtol=1e-5, atol=2e-6)
    pri = np.where(legal, np.exp(op - np.where(legal, op, -np.inf).max(1, keepdims=True)), 0)
    pri = pri / pri.sum(1, keepdims=True)
    np.testing.assert_allclose(out["priors"].cpu().numpy(), pri, rtol=1e-5, atol=1e-7)
    np.testing.assert_allclose(out["value"].cpu().numpy(), onet.support_to_scalar(ov, cfg.support_size)[:, 0], rtol=3e-4, atol=2e-4)
    act = rs.randint(A, size=B)
    out2 = net.recurrent_inference_fused(out["state"], torch.tensor(act, device=DEV), want_logits=True)
    ov2, or2, op2, os2 = onn.recurrent_inference(os_, act)
    np.testing.assert_allclose(out2["state"].cpu().numpy(), os2, rtol=1e-5, atol=3e-6)
    np.testing.assert_allclose(out2["reward_logits"].cpu().numpy(), or2, rtol=1e-5, atol=3e-6)
    np.testing.assert_allclose(out2["reward"].cpu().numpy(), onet.support_to_scalar(or2, cfg.support_size)[:, 0], rtol=3e-4, atol=2e-4)
    np.testing.assert_allclose(out2["priors"].cpu().numpy().sum(1), 1.0, atol=1e-6)
    # slot-addressed input (tree hidden-state pool): same result as the dense call
    H = cfg.encoding_size
    pool = torch.zeros((B, 3, H), device=DEV)
    slot = torch.tensor(rs.randint(3, size=B), dtype=torch.int32, device=DEV)
    pool[torch.arange(B, device=DEV), slot.long()] = out["state"]
    out3 = net.recurrent_inference_fused(pool, torch.tensor(act, device=DEV), in_slot=slot, in_row_stride=3 * H,
                                         slot_stride=H, state_out=pool, out_row_stride=3 * H, out_offset=0)
    assert torch.equal(out3["value"], out2["value"]) and torch.equal(out3["priors"], out2["priors"])


def test_codec_kernels():
    from muzero_hypermodel_b200 import models
    z = T.load("codec")
    s = models.support_to_scalar(torch.tensor(z["logits"], device=DEV), 10).cpu().numpy()
    np.testing.assert_allclose(s, z["scalars"], rtol=3e-4, atol=2e-4)
    # one-hot rows have an exact expectation -> the float32 transform is bit-exact; the centre row is excluded:
    # its expectation is a rounding residue of +-1e-20 whose SIGN (summation order) picks +-2.6e-5.
    rows = [r for r in range(256, 277) if r != 266]
    np.testing.assert_array_equal(s[rows], z["scalars"][rows])
    assert s[-2, 0] == 0
    sup = models.scalar_to_support(torch.tensor(z["x"], device=DEV), 10).cpu().numpy()
    np.testing.assert_allclose(sup, z["support"], rtol=0, atol=2e-6)
    # round trip: decode(encode(x)) == x for |h(x)| inside the support
    x = torch.linspace(-100, 100, 4001, device=DEV).reshape(1, -1)
    back = models.support_to_scalar(torch.log(models.scalar_to_support(x, 10)[0] + 1e-30), 10)
    np.testing.assert_allclose(back[:, 0].cpu().numpy(), x[0].cpu().numpy(), rtol=2e-3, atol=2e-3)


def test_cpu_tensor_is_refused():
    from muzero_hypermodel_b200 import models
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        models.support_to_scalar(torch.zeros(2, 21), 10)
