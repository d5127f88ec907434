#!/usr/bin/env python
"""Randomized parity of the two sibling gather ops against their numpy oracles (run on a GPU box):
    python tools/fuzz_siblings.py [trials=40] [seed=0]"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from iseg_b200.layers.dcn_v2 import dcnv2_sample  # noqa: E402
from iseg_b200.layers.deformable_attention import deform_attn_sample  # noqa: E402
from oracle import dcnv2_oracle as D2  # noqa: E402
from oracle import deform_attn_oracle as DA  # noqa: E402


def rel(a, b):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    return float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-30))


def run(fn, arrs, go, dt):
    t = [torch.from_numpy(a).to("cuda", dt).requires_grad_(True) for a in arrs]
    out = fn(*t)
    out.backward(torch.from_numpy(go).to("cuda", dt))
    return [out.detach().float().cpu().numpy()] + [v.grad.float().cpu().numpy() for v in t]


def main():
    trials = int(sys.argv[1]) if len(sys.argv) > 1 else 40
    rng = np.random.default_rng(int(sys.argv[2]) if len(sys.argv) > 2 else 0)
    fails = 0
    for i in range(trials):
        bf16 = rng.random() < 0.3
        dt, tol = (torch.bfloat16, 1e-2) if bf16 else (torch.float32, 1e-5)
        rnd = (lambda a: torch.from_numpy(a).bfloat16().float().numpy()) if bf16 else (lambda a: a)
        n, h, w = int(rng.integers(1, 4)), int(rng.integers(1, 40)), int(rng.integers(1, 40))
        # deformable attention
        heads, pts, c = int(rng.integers(1, 6)), int(rng.integers(1, 7)), int(rng.choice([1, 3, 4, 8, 20, 32, 40]))
        far = float(rng.choice([0.0, 2.0, 50.0]))
        v = rnd(rng.standard_normal((n, h, w, heads, c)).astype(np.float32))
        y = rnd(rng.uniform(-far, h - 1 + far, (n, h, w, heads, pts)).astype(np.float32))
        x = rnd(rng.uniform(-far, w - 1 + far, (n, h, w, heads, pts)).astype(np.float32))
        a = rnd(rng.uniform(-1, 1, (n, h, w, heads, pts)).astype(np.float32))
        go = rnd((rng.standard_normal((n, h, w, heads, c)) * float(rng.choice([1.0, 1e-4, 1e3]))).astype(np.float32))
        got = run(deform_attn_sample, (v, y, x, a), go, dt)
        ref = [DA.forward(v, y, x, a)] + list(DA.backward(v, y, x, a, go))
        bad = {k: rel(g, r) for k, g, r in zip(("out", "gv", "gy", "gx", "ga"), got, ref) if not rel(g, r) <= tol}
        if bad:
            fails += 1
            print("FAIL deform_attn", dict(n=n, h=h, w=w, heads=heads, pts=pts, c=c, far=far, bf16=bf16), bad, flush=True)
        # DCNv2 sampler
        k, c = int(rng.choice([3, 3, 5, 7])), int(rng.choice([1, 3, 4, 16, 24, 48]))
        xx = rnd(rng.standard_normal((n, h, w, c)).astype(np.float32))
        offs = rnd((rng.standard_normal((n, h, w, k * k, 2)) * float(rng.choice([0.5, 2.0, 20.0]))).astype(np.float32))
        m = rnd(rng.uniform(0, 1, (n, h, w, k * k)).astype(np.float32))
        go = rnd(rng.standard_normal((n, h, w, k * k, c)).astype(np.float32))
        got = run(lambda p, q, r: dcnv2_sample(p, q, r, k), (xx, offs, m), go, dt)
        ref = [D2.sample_forward(xx, offs, m, k, k)] + list(D2.sample_backward(xx, offs, m, go, k, k))
        bad = {kk: rel(g, r) for kk, g, r in zip(("out", "gx", "goff", "gm"), got, ref) if not rel(g, r) <= tol}
        if bad:
            fails += 1
            print("FAIL dcnv2", dict(n=n, h=h, w=w, c=c, k=k, bf16=bf16), bad, flush=True)
    print(f"sibling fuzz: {trials} trials each, {fails} failures")
    sys.exit(1 if fails else 0)


if __name__ == "__main__":
    main()
