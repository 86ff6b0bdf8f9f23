"""CPU: pins the autograd restatement of BCAgent.update (oracle/bc.py) to the literal formulation: which leaves the loss reaches
(`Policy`'s stop_gradient covers the image embeddings only) and finite differences of the loss."""
import numpy as np
import torch

from helpers import random_transitions


def _params(rng, cams, S=7, A=4):
    from serl_b200.agents.continuous.bc import bc_spec
    from serl_b200.params import ENC, init_trunk, lecun_normal, xavier_uniform
    spec, _ = bc_spec(cams, S, A)
    p = {}
    for l in spec:
        if l.path.endswith("kernel"):
            v = lecun_normal(rng, l.shape) if "/encoder_" in l.path else xavier_uniform(rng, l.shape)
        elif l.path.endswith("scale"):
            v = 1 + 0.1 * rng.standard_normal(l.shape)
        else:
            v = 0.05 * rng.standard_normal(l.shape)
        p[l.path] = torch.as_tensor(np.asarray(v, np.float32))
    for cam in cams:
        for k, v in init_trunk(rng).items():
            p[f"{ENC}/encoder_{cam}/pretrained_encoder/{k}"] = torch.as_tensor(v)
    return p


def test_bc_loss_reaches_mlp_heads_and_proprio_encoder_only_and_matches_finite_differences():
    from oracle import bc as OB
    cams = ("front",)
    rng = np.random.default_rng(0)
    params = _params(rng, cams)
    trs = random_transitions(rng, 3, cams)
    batch = {"observations": {"front": np.stack([t["observations"]["front"] for t in trs]), "state": np.stack([t["observations"]["state"] for t in trs])},
             "actions": np.stack([t["actions"] for t in trs]).astype(np.float32)}
    masks = {"front": rng.random((3, 4096)) < 0.9}
    key = np.array([0, 9], np.uint32)

    def run(p):
        opt = {"count": 0, "mu": {k: torch.zeros_like(v, dtype=torch.float64) for k, v in p.items() if "pretrained" not in k},
               "nu": {k: torch.zeros_like(v, dtype=torch.float64) for k, v in p.items() if "pretrained" not in k}}
        return OB.update(p, opt, key, cams, batch, dropout_masks=masks)

    newp, opt, new_rng, info, grads = run(params)
    for k, g in grads.items():
        if "/encoder_" in k:
            assert float(g.abs().max()) == 0.0, k                      # behind stop_gradient (common/encoding.py:48-49)
        else:
            assert float(g.abs().max()) > 0.0, k
    assert opt["count"] == 1 and not np.array_equal(new_rng, key)
    for path in ("modules_actor/network/Dense_0/kernel", "modules_actor/encoder/Dense_0/kernel", "modules_actor/Dense_1/bias"):
        flat = params[path].double().reshape(-1)
        for idx in rng.integers(0, flat.numel(), 2):
            h = 1e-5
            vals = []
            for sgn in (+1, -1):
                p2 = {k: v.double() for k, v in params.items()}
                t = p2[path].clone().reshape(-1)
                t[idx] += sgn * h
                p2[path] = t.reshape(params[path].shape)
                vals.append(run(p2)[3]["actor_loss"])
            fd = (vals[0] - vals[1]) / (2 * h)
            an = float(grads[path].reshape(-1)[idx])
            assert abs(fd - an) <= 1e-5 * max(abs(an), 1e-3) + 1e-7, (path, int(idx), fd, an)
    # one Adam step from zero moments moves every trained entry by lr (|m_hat / sqrt(v_hat)| = 1 where the gradient is non-zero)
    k = "modules_actor/network/Dense_1/kernel"
    moved = (newp[k] - params[k].double()).abs()
    assert abs(float(moved.max()) - 3e-4) < 1e-6
