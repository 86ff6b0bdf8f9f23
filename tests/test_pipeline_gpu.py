"""GPU: the cross-step pipeline of `update_critics` (sampler + frozen trunk of step i+1 next to the heads / Adam of step i, ping-pong
engines, CUDA-graph variants "W" / "P") against the serial path on the same handle sequence: crops and indices bit-exact, the key
chain identical, parameters equal up to the summation order of the GroupNorm statistics (shared-memory atomics in the conv epilogues)."""
import numpy as np
import pytest
import torch

from helpers import fake_env, random_transitions

pytestmark = pytest.mark.gpu


def _mk(cams, precision):
    from serl_b200.utils.launcher import make_drq_agent, make_replay_buffer
    env = fake_env(cams)
    rb = make_replay_buffer(env, capacity=300, type="memory_efficient_replay_buffer", image_keys=list(cams), seed=3)
    demo = make_replay_buffer(env, capacity=120, type="memory_efficient_replay_buffer", image_keys=list(cams), seed=4)
    trs = random_transitions(np.random.default_rng(1), 420, cams)
    for tr in trs[:300]:
        rb.insert(tr)
    for tr in trs[300:]:
        demo.insert(tr)
    agent = make_drq_agent(7, trs[0]["observations"], trs[0]["actions"], image_keys=cams, encoder_type="resnet-pretrained", precision=precision)
    return agent, rb, demo


@pytest.mark.parametrize("precision", ["fp16", "fp32"])
def test_pipelined_steps_equal_serial_steps(precision):
    from serl_b200.utils.train_utils import concat_batches
    cams, B = ("front", "wrist"), 32
    a_pipe, rb1, demo1 = _mk(cams, precision)
    a_ser, rb2, demo2 = _mk(cams, precision)
    a_pipe.pipeline_critic_steps = True
    its = [r.get_iterator(sample_args={"batch_size": B // 2, "pack_obs_and_next_obs": True}) for r in (rb1, demo1, rb2, demo2)]
    tol = 2e-3 if precision == "fp16" else 2e-5
    for step in range(7):                                            # W eager, P eager x2, P capture x2, P replay ...
        a_pipe, i1 = a_pipe.update_critics(concat_batches(next(its[0]), next(its[1]), axis=0))
        a_ser, i2 = a_ser.update_critics(concat_batches(next(its[2]), next(its[3]), axis=0))
        e1, e2 = a_pipe._last_engine, a_ser._engines[B]
        for cam in cams:
            assert torch.equal(e1.pix[cam], e2.pix[cam]), (step, cam)
        assert torch.equal(e1.idx, e2.idx) and torch.equal(e1.actions, e2.actions)
        np.testing.assert_array_equal(a_pipe.state.rng, a_ser.state.rng)
        assert a_pipe.state.step == a_ser.state.step == step + 1
        l1, l2 = float(i1["critic"]["critic_loss"]), float(i2["critic"]["critic_loss"])
        assert abs(l1 - l2) <= tol * max(abs(l2), 1e-6), (step, l1, l2)
        p1, p2 = a_pipe._store.params, a_ser._store.params
        assert float((p1 - p2).abs().max()) <= tol * float(p2.abs().max()), step
    # a call that does not continue the sequence (one handle skipped on both sides) restarts the pipeline and still agrees
    for it in its:
        next(it)
    a_pipe, i1 = a_pipe.update_critics(concat_batches(next(its[0]), next(its[1]), axis=0))
    a_ser, i2 = a_ser.update_critics(concat_batches(next(its[2]), next(its[3]), axis=0))
    assert torch.equal(a_pipe._last_engine.idx, a_ser._engines[B].idx)
    np.testing.assert_array_equal(a_pipe.state.rng, a_ser.state.rng)
    a_pipe.check_status(); a_ser.check_status()
