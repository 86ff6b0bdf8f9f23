"""Host-side cProfile of the e2e loop (insert -> sample -> update_critics -> loss readback)."""
import cProfile, pstats, sys, os, io
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np, torch
from helpers import fake_env, random_transitions
from serl_b200.utils.launcher import make_drq_agent, make_replay_buffer
import bench
cams = ("cam0",)
rb = make_replay_buffer(fake_env(cams), capacity=int(os.environ.get('CAP', 100000)), type="memory_efficient_replay_buffer", image_keys=list(cams), seed=1)
bench.fill_ring_synthetic(rb, 0)
rng = np.random.default_rng(0)
trs = random_transitions(rng, 8, cams, mean_ep=1000)
agent = make_drq_agent(42, trs[0]["observations"], trs[0]["actions"], image_keys=cams, encoder_type="resnet-pretrained", precision="fp16")
it = rb.get_iterator(sample_args={"batch_size": 256, "pack_obs_and_next_obs": True})
for _ in range(4):
    agent.update_critics(next(it))
torch.cuda.synchronize()
def loop(n):
    for s in range(n):
        rb.insert(trs[s % 8])
        _, info = agent.update_critics(next(it))
        float(info["critic"]["critic_loss"])
loop(5)
import time
torch.cuda.synchronize(); t0 = time.perf_counter(); loop(40); torch.cuda.synchronize(); print(f"e2e loop: {(time.perf_counter()-t0)/40*1e3:.2f} ms/step")
t0 = time.perf_counter()
for s in range(40): rb.insert(trs[s % 8])
print(f"insert only: {(time.perf_counter()-t0)/40*1e3:.3f} ms"); t0 = time.perf_counter(); rb.flush(); print(f"flush 40: {(time.perf_counter()-t0)*1e3:.3f} ms")
pr = cProfile.Profile(); pr.enable(); loop(40); pr.disable()
st = io.StringIO(); pstats.Stats(pr, stream=st).sort_stats("cumulative").print_stats(28); print(st.getvalue()[:6000])
