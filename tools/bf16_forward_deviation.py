"""Deviation of the default bf16 forward from the fp32 oracle on the bf16-rounded inputs (max-norm and rms), three
InternImage-T layer shapes.  usage: python tools/bf16_forward_deviation.py <label>   (GPU box; the oracle is the checker)"""
import sys
import numpy as np
import torch
sys.path.insert(0, '.')
from iseg_b200.layers.dcn_v3.op import dcnv3_op
from oracle import dcnv3_oracle as O
torch.manual_seed(0)
for (n, h, c, g) in [(2, 64, 128, 8), (2, 32, 256, 16), (1, 128, 64, 4)]:
    x = torch.randn(n, h, h, c).bfloat16(); off = (torch.randn(n, h, h, g * 18)).bfloat16()
    m = torch.softmax(torch.randn(n, h, h, g, 9), -1).reshape(n, h, h, g * 9).bfloat16()
    kw = dict(kernel_size=(3, 3), strides=(1, 1), padding="SAME", dilation_rate=(1, 1), groups=g, group_channels=16, offset_scale=1.0)
    ref = O.forward(x.float().numpy(), off.float().numpy(), m.float().numpy(), **kw)
    out = dcnv3_op(x.cuda(), off.cuda(), m.cuda(), (3, 3), (1, 1), "SAME", (1, 1), g, 16, 1.0).float().cpu().numpy()
    d = np.abs(out - ref)
    print(sys.argv[1], (n, h, c, g), "max|d|/max|ref| = %.3e  rms(d)/rms(ref) = %.3e" % (d.max() / np.abs(ref).max(), np.sqrt((d**2).mean()) / np.sqrt((ref**2).mean())))
