#!/usr/bin/env python
"""Time-boxed randomized parity fuzzing of the tiled DCNv3 path against the C oracle (run on a GPU box):

    python tools/fuzz_gpu.py [seconds=180] [seed=0]

Random image sizes up to 160 with wild aspect ratios, 1..12 groups of 16 or 32 channels, offset_scale in {0.5, 1, 2}, offset spread from 0
to far beyond every staged halo / scatter ring, fp32 and bf16, probability / logit / raw signed masks.  Consecutive
trials of equal workspace size reuse one zeroed-once workspace, so a path that leaves it dirty shows up as well.
Prints one line per failure and a summary; exit code 1 if anything failed.
"""
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import iseg_b200  # noqa: E402
from oracle import c_oracle  # noqa: E402
from oracle import dcnv3_oracle as O  # noqa: E402


def rel(a, b):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    return float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-30))


def main():
    budget = float(sys.argv[1]) if len(sys.argv) > 1 else 180.0
    rng = np.random.default_rng(int(sys.argv[2]) if len(sys.argv) > 2 else 0)
    t0, trials, fails = time.time(), 0, 0
    while time.time() - t0 < budget:
        trials += 1
        n = int(rng.integers(1, 4))
        h, w = int(rng.integers(3, 161)), int(rng.integers(3, 161))
        if rng.random() < 0.3:  # strongly non-square
            h, w = (int(rng.integers(3, 24)), int(rng.integers(100, 400))) if rng.random() < 0.5 else \
                   (int(rng.integers(100, 400)), int(rng.integers(3, 24)))
        g = int(rng.integers(1, 13))
        gc = int(rng.choice([16, 16, 32]))   # 32: InternImage-H, tiled kernels on half groups
        scale = float(rng.choice([0.5, 1.0, 1.0, 2.0]))
        sigma = float(rng.choice([0.0, 0.5, 1.0, 1.0, 3.0, 8.0]))
        bf16 = rng.random() < 0.35
        kind = str(rng.choice(["prob", "prob", "logits", "raw"]))
        x = rng.standard_normal((n, h, w, g * gc), dtype=np.float32)
        off = (sigma * rng.standard_normal((n, h, w, g * 18), dtype=np.float32)).astype(np.float32)
        if rng.random() < 0.3:
            off.reshape(-1)[::53] *= 40.0
        m = rng.standard_normal((n, h, w, g * 9), dtype=np.float32)
        go = rng.standard_normal((n, h, w, g * gc), dtype=np.float32) * float(rng.choice([1.0, 1e-5, 1e3]))
        if bf16:
            rnd = lambda a: torch.from_numpy(a).bfloat16().float().numpy()  # noqa: E731
            x, off, m, go = rnd(x), rnd(off), rnd(m), rnd(go)
        mask = O.mask_softmax(m, g) if kind != "raw" else m
        if bf16 and kind == "prob":
            mask = torch.from_numpy(mask).bfloat16().float().numpy()
        kw = dict(groups=g, group_channels=gc, offset_scale=scale)
        dt = torch.bfloat16 if bf16 else torch.float32
        tx, to = (torch.from_numpy(a).to("cuda", dt).requires_grad_() for a in (x, off))
        tm = torch.from_numpy(m if kind == "logits" else mask).to("cuda", dt).requires_grad_()
        blend = kind != "raw" and rng.random() < 0.3   # the centre-feature-scale blend fused around the op
        cs = ts = None
        if blend:
            cs = rng.uniform(-0.5, 1.5, (n, h, w, g)).astype(np.float32)
            if bf16:
                cs = torch.from_numpy(cs).bfloat16().float().numpy()
            ts = torch.from_numpy(cs).to("cuda", dt).requires_grad_()
            out = iseg_b200.dcnv3_op_center_scale(tx, to, tm, ts, [3, 3], [1, 1], "SAME", [1, 1], g, gc, scale,
                                                  mask_is_logits=kind == "logits")
        else:
            out = iseg_b200.dcnv3_op(tx, to, tm, [3, 3], [1, 1], "SAME", [1, 1], g, gc, scale, mask_is_logits=kind == "logits")
        out.backward(torch.from_numpy(go).to("cuda", dt))
        ref_out = c_oracle.forward(x, off, mask, **kw)
        go_core = go
        if blend:  # out = core * (1 - s) + x * s; the core sees grad_out * (1 - s), x also grad_out * s, d s = sum go * (x - core)
            s16 = np.repeat(cs, gc, axis=-1)
            rs = (go * (x - ref_out)).reshape(n, h, w, g, gc).sum(-1)
            go_core = (go * (np.float32(1) - s16)).astype(np.float32)
            ref_out = ref_out * (np.float32(1) - s16) + x * s16
        _, roff, rm = c_oracle.backward(x, off, mask, go_core, **kw)
        rx, _, _ = O.backward(x, off, mask, go_core, accumulate=np.float64, **kw) if h * w * n * g < 60000 else \
            c_oracle.backward(x, off, mask, go_core, **kw)
        if blend:
            rx = rx + go * s16
        if kind == "logits":  # softmax Jacobian
            mm, gg = mask.reshape(n, h, w, g, 9), rm.reshape(n, h, w, g, 9)
            rm = (mm * (gg - (mm * gg).sum(-1, keepdims=True))).reshape(n, h, w, g * 9)
        tol = 1e-2 if bf16 else (1e-5 if kind != "raw" else 3e-5)
        errs = {"out": rel(out.detach().float().cpu().numpy(), ref_out), "gx": rel(tx.grad.float().cpu().numpy(), rx),
                "goff": rel(to.grad.float().cpu().numpy(), roff), "gm": rel(tm.grad.float().cpu().numpy(), rm)}
        if blend:
            errs["gs"] = rel(ts.grad.float().cpu().numpy(), rs)
        bad = {k: v for k, v in errs.items() if not v <= tol}
        if bad:
            fails += 1
            print("FAIL", dict(n=n, h=h, w=w, g=g, scale=scale, sigma=sigma, bf16=bf16, mask=kind, blend=blend), bad, flush=True)
    print(f"fuzz: {trials} trials, {fails} failures, {time.time() - t0:.0f} s")
    sys.exit(1 if fails else 0)


if __name__ == "__main__":
    main()
