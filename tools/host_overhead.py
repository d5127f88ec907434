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
# bare C ABI (raw pointers, prebuilt parameter block): what the library itself costs on the host per call
import ctypes
p=_cabi.make_params(xd.shape,(h,w),*cfg[:4],g,16,1.0,_cabi.F32,_cabi.FLAG_WORKSPACE_ZEROED)
out=torch.empty_like(xd); gx=torch.empty_like(xd); goff=torch.empty_like(od); gm=torch.empty_like(md)
wsb=int(_cabi.lib.dcnv3_backward_workspace_bytes(ctypes.byref(p))); ws=torch.zeros(wsb,dtype=torch.uint8,device="cuda")
st=ctypes.c_void_p(torch.cuda.current_stream().cuda_stream); vp=lambda t: ctypes.c_void_p(t.data_ptr())
fa=[vp(t) for t in (xd,od,md,out)]; ba=[vp(t) for t in (xd,od,md,go,gx,goff,gm,ws)]
torch.cuda.synchronize(); t0=time.perf_counter()
for _ in range(N): _cabi.lib.dcnv3_forward(*fa,ctypes.byref(p),st)
torch.cuda.synchronize(); print("dcnv3_forward (raw pointers): %.1f us"%((time.perf_counter()-t0)/N*1e6))
t0=time.perf_counter()
for _ in range(N): _cabi.lib.dcnv3_backward(*ba,wsb,ctypes.byref(p),st)
torch.cuda.synchronize(); print("dcnv3_backward (raw pointers): %.1f us"%((time.perf_counter()-t0)/N*1e6))
