"""GPU: BCAgent (serl_b200/agents/continuous/bc.py) against the CPU restatement of the reference's
BCAgent.update (oracle/bc.py): loss, mse, every gradient leaf (zero for the image heads behind stop_gradient, live for the
proprio encoder), parameters after Adam, the key chain, and sample_actions.  fp32 build: 1e-5 class bars."""
import numpy as np
import pytest
import torch

from helpers import random_transitions, rel_err

pytestmark = pytest.mark.gpu


def _flat(tree, prefix=""):
    out = {}
    for k, v in tree.items():
        p = f"{prefix}/{k}" if prefix else k
        out.update(_flat(v, p)) if isinstance(v, dict) else out.__setitem__(p, v)
    return out


@pytest.mark.parametrize("cams", [("front",), ("front", "wrist")])
def test_bc_update_matches_oracle(cams):
    from oracle import bc as OB
    from oracle import drq as O
    from serl_b200.utils.launcher import make_bc_agent
    rng = np.random.default_rng(0)
    B, A = 12, 4
    trs = random_transitions(rng, B, cams)
    agent = make_bc_agent(3, trs[0]["observations"], trs[0]["actions"], image_keys=cams, encoder_type="resnet-pretrained")
    g = torch.Generator(device="cuda").manual_seed(1)               # biases off zero so that every path is exercised
    agent._params.add_(torch.randn(agent._n, device="cuda", generator=g) * 0.05)
    batch = {"observations": {**{c: np.stack([t["observations"][c] for t in trs]) for c in cams},
                              "state": np.stack([t["observations"]["state"] for t in trs])},
             "actions": np.stack([t["actions"] for t in trs]).astype(np.float32)}
    opt = None
    for step in range(2):
        params = {k: torch.as_tensor(np.asarray(v)) for k, v in _flat(agent.state.params).items()}
        if opt is None:
            opt = {"count": 0, "mu": {k: torch.zeros_like(v, dtype=torch.float64) for k, v in params.items() if "pretrained_encoder" not in k},
                   "nu": {k: torch.zeros_like(v, dtype=torch.float64) for k, v in params.items() if "pretrained_encoder" not in k}}
        rng0 = agent.state.rng
        agent, info = agent.update(batch)
        newp, opt, new_rng, oinfo, grads = OB.update(params, opt, rng0, cams, batch)
        assert abs(float(info["actor_loss"]) - oinfo["actor_loss"]) <= 1e-5 * max(abs(oinfo["actor_loss"]), 1.0)
        assert abs(float(info["mse"]) - oinfo["mse"]) <= 1e-5 * max(abs(oinfo["mse"]), 1.0)
        np.testing.assert_array_equal(agent.state.rng, new_rng)
        for l in agent._spec:
            got = agent._grad[l.offset:l.offset + l.size].view(l.shape).cpu().numpy()
            ref = grads[l.path].numpy()
            if "/encoder_" in l.path:                               # image heads: behind stop_gradient (encoding.py:48-49)
                assert np.abs(ref).max() == 0 and np.abs(got).max() == 0, l.path
            else:
                assert np.abs(ref).max() > 0, l.path
                assert np.abs(got - ref).max() <= 2e-4 * np.abs(ref).max(), l.path
        now = _flat(agent.state.params)
        lr = agent.learning_rate
        for l in agent._spec:
            ref, got = newp[l.path].numpy(), np.asarray(now[l.path])
            gmag = np.abs(grads[l.path].numpy())
            noisy = gmag < 2e-2 * max(gmag.max(), 1e-30)             # Adam normalises by |g|: entries at noise level move by up to ~lr either way
            allow = 1e-5 * max(np.abs(ref).max(), 1e-3) + lr * np.where(noisy, 2.2, 5e-3)
            assert (np.abs(got - ref) <= allow).all(), (l.path, np.abs(got - ref).max())
    assert agent.state.step == 2


def test_bc_sample_actions_and_debug_metrics():
    from oracle import bc as OB
    from oracle import drq as O
    from oracle.jax_prng import normal
    from serl_b200.utils.launcher import make_bc_agent
    cams = ("front",)
    rng = np.random.default_rng(4)
    trs = random_transitions(rng, 5, cams)
    agent = make_bc_agent(9, trs[0]["observations"], trs[0]["actions"], image_keys=cams, encoder_type="resnet-pretrained")
    obs = {"front": np.stack([t["observations"]["front"] for t in trs]), "state": np.stack([t["observations"]["state"] for t in trs])}
    params = {k: torch.as_tensor(np.asarray(v)).double() for k, v in _flat(agent.state.params).items()}
    img = torch.as_tensor(obs["front"])
    b, t, h, w, c = img.shape
    feats = {"front": O.trunk_forward(params, "front", img.permute(0, 2, 3, 1, 4).reshape(b, h, w, t * c), torch.float64)}
    mu, sd = OB.bc_forward(params, cams, feats, torch.as_tensor(obs["state"]).double(), None)
    a = agent.sample_actions(obs, argmax=True)
    assert rel_err(a, mu.numpy()) < 1e-5
    seed = np.array([0, 11], np.uint32)
    s = agent.sample_actions(obs, seed=seed)
    eps = normal(seed, (5, 4))
    assert rel_err(s, (mu + sd * torch.as_tensor(eps)).numpy()) < 1e-5
    one = agent.sample_actions({k: v[0] for k, v in obs.items()}, argmax=True)
    assert one.shape == (4,) and rel_err(one, mu.numpy()[0]) < 1e-5
    m = agent.get_debug_metrics({"observations": obs, "actions": np.zeros((5, 4), np.float32)})
    assert rel_err(m["mse"].cpu().numpy(), (mu ** 2).sum(-1).numpy()) < 1e-5
