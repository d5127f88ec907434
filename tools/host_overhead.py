"""Host-side cost of one small dcnv3_op forward + backward through torch autograd vs the bare ctypes calls
(run on a GPU box): python tools/host_overhead.py"""
import time, torch, sys, os
sys.path.insert(0, os.getcwd())
import iseg_b200
from iseg_b200 import _cabi
n,h,w,g=2,16,16,4
x=torch.randn(n,h,w,g*16,device="cuda",requires_grad=True); off=torch.randn(n,h,w,g*18,device="cuda",requires_grad=True)
m=torch.softmax(torch.randn(n,h,w,g,9,device="cuda"),-1).reshape(n,h,w,g*9).requires_grad_()
go=torch.randn(n,h,w,g*16,device="cuda")
def step():
    out=iseg_b200.dcnv3_op(x,off,m,[3,3],[1,1],"SAME",[1,1],g,16,1.0)
    out.backward(go)
for _ in range(20): step()
torch.cuda.synchronize(); t0=time.perf_counter()
N=500
for _ in range(N): step()
torch.cuda.synchronize(); dt=(time.perf_counter()-t0)/N
print("dcnv3_op fwd+bwd eager (tiny shape): %.1f us per call pair"%(dt*1e6))
cfg=((3,3),(1,1),(1,1),(1,1),g,16,1.0)
xd,od,md=x.detach(),off.detach(),m.detach()
torch.cuda.synchronize(); t0=time.perf_counter()
for _ in range(N): _cabi.forward(xd,od,md,*cfg)
torch.cuda.synchronize(); print("_cabi.forward: %.1f us"%((time.perf_counter()-t0)/N*1e6))
t0=time.perf_counter()
for _ in range(N): _cabi.backward(xd,od,md,go,*cfg)
torch.cuda.synchronize(); print("_cabi.backward: %.1f us"%((time.perf_counter()-t0)/N*1e6))
