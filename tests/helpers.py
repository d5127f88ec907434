"""Shared helpers for the parity tests."""
import glob
import os

import numpy as np

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def golden_op_cases(bf16=False):
    """fp32 / fp64 fixtures by default; bf16=True: the reference run on bfloat16 tensors (op_bf16_*)."""
    names = sorted(os.path.basename(p)[3:-4] for p in glob.glob(os.path.join(GOLDEN, "op_*.npz")))
    return [n for n in names if n.startswith("bf16_") == bf16]


def load_op_case(name):
    z = np.load(os.path.join(GOLDEN, f"op_{name}.npz"))
    if "bf16_bits" in z.files:  # bfloat16 fixtures are stored as uint16 bit patterns: widen to float32
        z = {k: ((z[k].astype(np.uint32) << 16).view(np.float32) if z[k].dtype == np.uint16 else z[k]) for k in z.files}
    kw = dict(kernel_size=tuple(int(v) for v in z["kernel_size"]),
              strides=tuple(int(v) for v in z["strides"]),
              padding=str(z["padding"]),
              dilation_rate=tuple(int(v) for v in z["dilation_rate"]),
              groups=int(z["groups"]), group_channels=int(z["group_channels"]),
              offset_scale=float(z["offset_scale"]))
    return z, kw


def rel_err(a, b):
    """max|a-b| / max|b| -- the parity metric of BASELINE.md section 5."""
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    return float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-30))


def make_inputs(n, h, w, g, gc, sigma=1.0, seed=0, dtype=np.float32, p=9):
    """Synthetic inputs of SURVEY.md section 8(d): x~N(0,1), offset~N(0,sigma^2), mask=softmax_P(N(0,1)),
    grad_out~N(0,1)."""
    rng = np.random.default_rng(seed)
    x = rng.standard_normal((n, h, w, g * gc), dtype=np.float32)
    offset = (sigma * rng.standard_normal((n, h, w, g * p * 2), dtype=np.float32)).astype(np.float32)
    z = rng.standard_normal((n, h, w, g, p), dtype=np.float32)
    e = np.exp(z - z.max(-1, keepdims=True))
    mask = (e / e.sum(-1, keepdims=True)).reshape(n, h, w, g * p).astype(np.float32)
    grad_out = rng.standard_normal((n, h, w, g * gc), dtype=np.float32)
    return tuple(a.astype(dtype) for a in (x, offset, mask, grad_out))
