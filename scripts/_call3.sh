set -u
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_tgemm_gpu.py -m gpu -q --timeout=300 -p no:cacheprovider 2>&1 | grep -E "^(FAILED|PASSED|ERROR)|passed|failed|AssertionError: assert" | head -40
python - <<'PY'
import numpy as np, torch
from serl_b200 import ops
rng=np.random.default_rng(0)
M,K,N=128,32,256
x=torch.as_tensor(rng.standard_normal((M,K)).astype(np.float32)).cuda()
w=torch.as_tensor(rng.standard_normal((K,N)).astype(np.float32)).cuda()
out=torch.zeros(M,N,device='cuda')
p=ops.tgemm_problem(x.data_ptr(), w.data_ptr(), sAm=K, sAk=1, sBk=N, sBn=1, C_=out.data_ptr(), ldc=N)
ops.tgemm(None,[p],M,N,K)
ref=x.double()@w.double()
print('fwd X@W  max|out|',out.abs().max().item(),'max|ref|',ref.abs().max().item(),'err',(out-ref).abs().max().item())
# which structure? compare with candidates
wt=w.t().contiguous()
out2=torch.zeros(M,N,device='cuda')
p=ops.tgemm_problem(x.data_ptr(), wt.data_ptr(), sAm=K, sAk=1, sBk=1, sBn=K, C_=out2.data_ptr(), ldc=N)
ops.tgemm(None,[p],M,N,K)
print('K-major B: err',(out2-ref).abs().max().item())
xt=x.t().contiguous()
out3=torch.zeros(M,N,device='cuda')
p=ops.tgemm_problem(xt.data_ptr(), wt.data_ptr(), sAm=1, sAk=M, sBk=1, sBn=K, C_=out3.data_ptr(), ldc=N)
ops.tgemm(None,[p],M,N,K)
print('MN-major A, K-major B: err',(out3-ref).abs().max().item())
torch.set_printoptions(precision=3, linewidth=200)
print(out[:4,:8]); print(ref[:4,:8])
# identity probes for the MN-major B layout: x = e_k rows pick rows of w
K=32
for kk in (0,1,8,9):
    x=torch.zeros(M,K,device='cuda'); x[:,kk]=1
    out=torch.zeros(M,N,device='cuda')
    p=ops.tgemm_problem(x.data_ptr(), w.data_ptr(), sAm=K, sAk=1, sBk=N, sBn=1, C_=out.data_ptr(), ldc=N)
    ops.tgemm(None,[p],M,N,K)
    # find which w element each output column equals
    row=out[0]
    hits=[]
    for n in (0,1,4,31,32,33,255):
        m=(w-row[n]).abs()<1e-3
        idx=m.nonzero()
        hits.append((n, idx[:3].tolist()))
    print('k',kk,hits)
PY
