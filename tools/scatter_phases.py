"""Per-CTA phase clocks of bwd_scatter_kernel (library built with -DDCNV3_SCATTER_PROFILE).
usage: python tools/scatter_phases.py LIB H W C G [batch] [dtype]"""
import ctypes
import os
import sys

import numpy as np
import torch

lib_path = sys.argv[1]
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from iseg_b200 import _cabi as cabi  # noqa: E402

dbg = ctypes.CDLL(lib_path)
h, w, c, g = (int(v) for v in sys.argv[2:6])
batch = int(sys.argv[6]) if len(sys.argv) > 6 else 16
dtype = sys.argv[7] if len(sys.argv) > 7 else "f32"
tdt = torch.float32 if dtype == "f32" else torch.bfloat16
gen = torch.Generator(device="cuda").manual_seed(0)
r = lambda *s: torch.randn(*s, device="cuda", generator=gen)  # noqa: E731
x = r(batch, h, w, c).to(tdt)
off = r(batch, h, w, g * 18).to(tdt)
mask = torch.softmax(r(batch, h, w, g, 9), -1).reshape(batch, h, w, g * 9).to(tdt)
go = r(batch, h, w, c).to(tdt)
gx, goff, gm = torch.empty_like(x), torch.empty_like(off), torch.empty_like(mask)
p = cabi.make_params(x.shape, (h, w), (3, 3), (1, 1), (1, 1), (1, 1), g, c // g, 1.0,
                     cabi.F32 if dtype == "f32" else cabi.BF16, cabi.FLAG_WORKSPACE_ZEROED)
for name in ("dcnv3_backward_workspace_bytes", "dcnv3_backward"):
    getattr(dbg, name).restype = getattr(cabi.lib, name).restype
    getattr(dbg, name).argtypes = getattr(cabi.lib, name).argtypes
wsb = int(dbg.dcnv3_backward_workspace_bytes(ctypes.byref(p)))
ws = torch.zeros(wsb, dtype=torch.uint8, device="cuda")
st = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
vp = lambda t: ctypes.c_void_p(t.data_ptr())  # noqa: E731
prof = torch.zeros(1 << 16, 16, dtype=torch.int64, device="cuda")
for i in range(3):
    if i == 2:
        assert dbg.dcnv3_debug_scatter_profile(ctypes.c_void_p(prof.data_ptr())) == 0
    rc = dbg.dcnv3_backward(vp(x), vp(off), vp(mask), vp(go), vp(gx), vp(goff), vp(gm), vp(ws), wsb, ctypes.byref(p), st)
    assert rc == 0, rc
    torch.cuda.synchronize()
a = prof.cpu().numpy()
a = a[a[:, 0] != 0]
d = lambda i, j: (a[:, j] - a[:, i]).astype(np.float64)  # noqa: E731
print(f"{h}x{w} C{c} G{g} {dtype}: {len(a)} CTAs on {len(set(a[:, 8]))} SMs; clocks, mean (min..max)")
for name, v in (("zero-init", d(0, 1)), ("pdl wait", d(1, 2)), ("walk (to barrier)", d(2, 3)), ("first warp done", a[:, 6] - a[:, 2]),
                ("last warp done", a[:, 7] - a[:, 2]), ("flush", d(3, 4)), ("whole CTA", d(0, 4))):
    v = np.asarray(v, dtype=np.float64)
    print(f"  {name:20s} {v.mean():10.0f}  ({v.min():.0f} .. {v.max():.0f})")
per_sm = {}
for row in a:
    per_sm.setdefault(int(row[8]), []).append((row[0], row[4]))
busy = np.array([sum(e - s for s, e in v) for v in per_sm.values()], dtype=np.float64)
span = np.array([max(e for s, e in v) - min(s for s, e in v) for v in per_sm.values()], dtype=np.float64)
print(f"  per SM: CTAs {np.mean([len(v) for v in per_sm.values()]):.2f}, busy {busy.mean():.0f}, span {span.mean():.0f} (max {span.max():.0f})")
