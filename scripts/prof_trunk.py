"""Runs the frozen trunk alone (profiling target for ncu)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np, torch
from helpers import random_transitions
from serl_b200.utils.launcher import make_drq_agent

prec = sys.argv[1] if len(sys.argv) > 1 else "fp16"
B = int(sys.argv[2]) if len(sys.argv) > 2 else 256
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 3
cams = ("cam0",)
tr = random_transitions(np.random.default_rng(0), 1, cams)[0]
agent = make_drq_agent(42, tr["observations"], tr["actions"], image_keys=cams, encoder_type="resnet-pretrained", precision=prec)
eng = agent._engine(B)
eng.pix["cam0"].copy_(torch.randint(0, 256, eng.pix["cam0"].shape, dtype=torch.uint8, device="cuda"))
for _ in range(reps):
    eng.trunk_forward("cam0", eng.pix["cam0"], eng.feats["cam0"])
torch.cuda.synchronize()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record()
for _ in range(reps):
    eng.trunk_forward("cam0", eng.pix["cam0"], eng.feats["cam0"])
b.record(); torch.cuda.synchronize()
print(f"trunk {prec} N={2*B}: {a.elapsed_time(b)/reps:.3f} ms per pass -> {2*B*0.5804/1e3/(a.elapsed_time(b)/reps/1e3):.1f} TFLOP/s")
