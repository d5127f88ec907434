"""Runs forward+backward of one DCNv3 layer shape a few times (for ncu / quick timing).
usage: python tools/prof_one.py H W C G [batch] [dtype] [reps]"""
import ctypes
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from iseg_b200 import _cabi as cabi  # noqa: E402

h, w, c, g = (int(v) for v in sys.argv[1:5])
batch = int(sys.argv[5]) if len(sys.argv) > 5 else 16
dtype = sys.argv[6] if len(sys.argv) > 6 else "f32"
reps = int(sys.argv[7]) if len(sys.argv) > 7 else 3
sigma = float(os.environ.get("OFFSET_SIGMA", "1.0"))
scale = float(os.environ.get("OFFSET_SCALE", "1.0"))
tdt = torch.float32 if dtype == "f32" else torch.bfloat16
gen = torch.Generator(device="cuda").manual_seed(0)
r = lambda *s: torch.randn(*s, device="cuda", generator=gen)  # noqa: E731
x = r(batch, h, w, c).to(tdt)
off = (sigma * r(batch, h, w, g * 18)).to(tdt)
mask = torch.softmax(r(batch, h, w, g, 9), -1).reshape(batch, h, w, g * 9).to(tdt)
go = r(batch, h, w, c).to(tdt)
out, gx, goff, gm = torch.empty_like(x), torch.empty_like(x), torch.empty_like(off), torch.empty_like(mask)
p = cabi.make_params(x.shape, (h, w), (3, 3), (1, 1), (1, 1), (1, 1), g, c // g, scale,
                     cabi.F32 if dtype == "f32" else cabi.BF16, cabi.FLAG_WORKSPACE_ZEROED)
wsb = int(cabi.lib.dcnv3_backward_workspace_bytes(ctypes.byref(p)))
ws = torch.zeros(wsb, dtype=torch.uint8, device="cuda")
st = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
vp = lambda t: ctypes.c_void_p(t.data_ptr())  # noqa: E731
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
tf, tb = [], []
for i in range(reps):
    flush.zero_()
    e = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
    e[0].record()
    cabi.check(cabi.lib.dcnv3_forward(vp(x), vp(off), vp(mask), vp(out), ctypes.byref(p), st))
    e[1].record()
    cabi.check(cabi.lib.dcnv3_backward(vp(x), vp(off), vp(mask), vp(go), vp(gx), vp(goff), vp(gm), vp(ws), wsb,
                                       ctypes.byref(p), st))
    e[2].record()
    torch.cuda.synchronize()
    tf.append(e[0].elapsed_time(e[1]) * 1e3)
    tb.append(e[1].elapsed_time(e[2]) * 1e3)
print(f"{h}x{w} C{c} G{g} b{batch} {dtype} sigma={sigma} scale={scale}: fwd {min(tf):.1f} us  bwd {min(tb):.1f} us")
