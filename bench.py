#!/usr/bin/env python
"""Headline benchmark: DCNv3 core-op forward+backward over the 30 DCNv3 layers of InternImage-T at a
512x512 crop, batch 16 per GPU (BASELINE.json configs[1]), sampled-points/s and achieved HBM GB/s.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--dtype f32|bf16]

One step = forward of layers 1..30 (stage shapes 128^2xC64/G4 x4, 64^2xC128/G8 x4, 32^2xC256/G16 x18,
16^2xC512/G32 x4; reference backbones/intern_image/intern_image.py:137-150) followed by backward of
layers 30..1, each layer with its own tensors (working set ~5.4 GB fp32 >> 126 MB L2).
A sampled point is one (n,h,w,g,p) tuple: 103.8 M per step per GPU.

Prints ONE JSON line (contract in the task statement; `roofline` and `cpu_baseline` added).
`--impl reference` times the CPU restatement of the reference path (oracle/dcnv3_oracle.c, all host
threads) on a bounded sample; TensorFlow, which the reference needs, is not installable here.
"""
import argparse
import ctypes
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

STAGES = [  # H, W, C, G, depth   (InternImage-T @512: stem /4, each downsample /2 and x2 channels)
    (128, 128, 64, 4, 4),
    (64, 64, 128, 8, 4),
    (32, 32, 256, 16, 18),
    (16, 16, 512, 32, 4),
]
BATCH = 16
P = 9
GC = 16
METRIC = "dcnv3_fwd_bwd_sampled_points_per_sec"
UNIT = "points/s"


def points_per_step(batch=BATCH):
    return sum(batch * h * w * g * P * d for h, w, _, g, d in STAGES)


def algo_bytes(h, w, c, g, esize, batch=BATCH):
    """fwd (2C+3GP)*b, bwd (3C+6GP)*b per output pixel."""
    px = batch * h * w
    return px * (2 * c + 3 * g * P) * esize, px * (3 * c + 6 * g * P) * esize


def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.isfile(path):
        return float(json.load(open(path))["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


# ------------------------------------------------------------------------------------------------
class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled during the timed region."""
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.rows, self.proc, self.index = [], None, index

    def __enter__(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                 "-lms", "50"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None
        return self

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def __exit__(self, *exc):
        if self.proc:
            time.sleep(0.15)
            self.proc.terminate()
            self.thread.join(timeout=2)

    def summary(self):
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[0])), mx.append(float(r[1]))
            except (ValueError, IndexError):
                continue
            for nm, v in zip(names, r[2:6]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2], "sm_max_mhz": max(mx), "reasons": sorted(reasons), "samples": len(sm)}


# ------------------------------------------------------------------------------------------------
def cpu_reference_run(steps, warmup, sample_batch=1, budget_s=None):
    """The reference path on the host CPU: C restatement, all threads, batch `sample_batch` of each of
    the 30 layers, forward + backward.  Returns (points/s, info)."""
    import numpy as np
    from oracle import c_oracle

    c_oracle.lib(native=True)
    # every host thread this process may use (torchrun exports OMP_NUM_THREADS=1; the CPU arm ignores it)
    try:
        threads = len(os.sched_getaffinity(0))
    except AttributeError:
        threads = os.cpu_count() or 1
    rng = np.random.default_rng(0)
    data = []
    for h, w, c, g, depth in STAGES:
        x = rng.standard_normal((sample_batch, h, w, c), dtype=np.float32)
        off = rng.standard_normal((sample_batch, h, w, g * P * 2), dtype=np.float32)
        z = rng.standard_normal((sample_batch, h, w, g, P), dtype=np.float32)
        e = np.exp(z - z.max(-1, keepdims=True))
        m = (e / e.sum(-1, keepdims=True)).reshape(sample_batch, h, w, g * P).astype(np.float32)
        go = rng.standard_normal((sample_batch, h, w, c), dtype=np.float32)
        data.append((x, off, m, go, g, depth))

    def one_step():
        for x, off, m, go, g, depth in data:
            for _ in range(depth):
                c_oracle.forward(x, off, m, groups=g, group_channels=GC, nthreads=threads)
        for x, off, m, go, g, depth in reversed(data):
            for _ in range(depth):
                c_oracle.backward(x, off, m, go, groups=g, group_channels=GC, nthreads=threads)

    for _ in range(warmup):
        one_step()
    t0 = time.perf_counter()
    done = 0
    for _ in range(steps):
        one_step()
        done += 1
        if budget_s is not None and time.perf_counter() - t0 > budget_s:
            break
    dt = time.perf_counter() - t0
    pts = points_per_step(sample_batch) * done
    info = {"kind": "port", "cores": threads,
            "sample": f"batch {sample_batch} of each of the 30 InternImage-T DCNv3 layers, fwd+bwd, fp32, "
                      f"{done} step(s) of {points_per_step(sample_batch)} points; oracle/dcnv3_oracle.c -O3 "
                      f"-march=native OpenMP (TensorFlow reference not installable)"}
    return pts / dt, dt / done * 1e3, info


def run_reference(args, rank, world):
    if rank != 0:
        return
    value, ms, info = cpu_reference_run(args.steps, max(args.warmup, 1), sample_batch=1)
    info["value"] = value
    info["unit"] = UNIT
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_config(args, "f32", world, launch="cpu"),
        "cpu_baseline": info,
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


def workload_config(args, dtype, world, launch):
    return {
        "workload": "InternImage-T DCNv3 core op fwd+bwd, all 30 layers (stages 128^2xC64/G4 x4, 64^2xC128/G8 x4, "
                    "32^2xC256/G16 x18, 16^2xC512/G32 x4), 512x512 crop",
        "batch_per_gpu": BATCH, "global_batch": BATCH * world, "kernel": "3x3 s1 d1 SAME", "group_channels": GC,
        "offset_sigma": 1.0, "parallelism": f"dp{world} (whole images sharded, no collective on the op)",
        "launch": launch,
        "l2": "each layer owns its tensors (working set >> 126 MB L2); fwd 1..30 then bwd 30..1",
    }


# ------------------------------------------------------------------------------------------------
class Layer:
    def __init__(self, torch, cabi, h, w, c, g, dtype, seed, device):
        gen = torch.Generator(device=device).manual_seed(seed)
        r = lambda *s: torch.randn(*s, device=device, generator=gen)  # noqa: E731
        tdt = torch.float32 if dtype == "f32" else torch.bfloat16
        self.x = r(BATCH, h, w, c).to(tdt)
        self.off = r(BATCH, h, w, g * P * 2).to(tdt)
        self.mask = torch.softmax(r(BATCH, h, w, g, P), -1).reshape(BATCH, h, w, g * P).to(tdt)
        self.go = r(BATCH, h, w, c).to(tdt)
        self.out = torch.empty_like(self.x)
        self.gx, self.goff, self.gm = torch.empty_like(self.x), torch.empty_like(self.off), torch.empty_like(self.mask)
        self.p = cabi.make_params(self.x.shape, (h, w), (3, 3), (1, 1), (1, 1), (1, 1), g, GC, 1.0,
                                  cabi.F32 if dtype == "f32" else cabi.BF16)
        self.pref = ctypes.byref(self.p)
        self.ws_bytes = int(cabi.lib.dcnv3_backward_workspace_bytes(self.pref))
        self.ws = torch.zeros(max(self.ws_bytes, 256), dtype=torch.uint8, device=device)
        self.p.flags |= cabi.FLAG_WORKSPACE_ZEROED  # zeroed once; dcnv3_backward leaves it zeroed
        self.shape = (h, w, c, g)
        self.fwd_args = [ctypes.c_void_p(t.data_ptr()) for t in (self.x, self.off, self.mask, self.out)]
        self.bwd_args = [ctypes.c_void_p(t.data_ptr()) for t in
                         (self.x, self.off, self.mask, self.go, self.gx, self.goff, self.gm, self.ws)]


def run_ours(args, rank, world, local_rank):
    import torch

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device (the DCNv3 path has no CPU fallback)")
    torch.cuda.set_device(local_rank)
    device = torch.device("cuda", local_rank)
    dist = None
    if world > 1:
        # keep NCCL's banner / debug output off stdout: rank 0 prints ONE JSON line there
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
        import torch.distributed as dist_mod
        dist = dist_mod
        dist.init_process_group("nccl", device_id=device)
    from iseg_b200 import _cabi as cabi

    lib = cabi.lib

    def build_layers(dtype):
        layers, seed = [], 1000 * rank
        for (h, w, c, g, depth) in STAGES:
            for _ in range(depth):
                layers.append(Layer(torch, cabi, h, w, c, g, dtype, seed, device))
                seed += 1
        return layers

    def make_step(layers, stream_ptr):
        fwd, bwd, chk = lib.dcnv3_forward, lib.dcnv3_backward, cabi.check

        def step():
            for l in layers:
                chk(fwd(*l.fwd_args, l.pref, stream_ptr))
            for l in reversed(layers):
                chk(bwd(*l.bwd_args, l.ws_bytes, l.pref, stream_ptr))
        return step

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps, warmup, sampler=None):
        for _ in range(warmup):
            fn()
        barrier()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        n0 = cabi.launch_count()
        if sampler:
            sampler.__enter__()
        a.record()
        for _ in range(steps):
            fn()
        b.record()
        barrier()
        if sampler:
            sampler.__exit__()
        ms = a.elapsed_time(b)
        if dist is not None:
            t = torch.tensor([ms], device=device)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms, cabi.launch_count() - n0

    def measure(dtype, sampler=None):
        layers = build_layers(dtype)
        stream = torch.cuda.Stream(device=device)
        res = {}
        with torch.cuda.stream(stream):
            sp = ctypes.c_void_p(stream.cuda_stream)
            step = make_step(layers, sp)
            launch = "direct"
            fn = step
            launches_per_step = None
            if args.graph:
                step()  # warm (lazy module load must not happen inside capture)
                torch.cuda.synchronize()
                n0 = cabi.launch_count()
                graph = torch.cuda.CUDAGraph()
                with torch.cuda.graph(graph, stream=stream):
                    step()
                launches_per_step = cabi.launch_count() - n0
                fn, launch = graph.replay, "cuda_graph"
            ms, launched = timed(fn, args.steps, args.warmup, sampler)
            if launches_per_step is not None:
                launched = launches_per_step * args.steps
            res.update(ms_total=ms, launches=launched, launch=launch)
            # ---- per-kernel durations (CUDA events on the launching stream, recorded around every kernel:
            #      forward from here, the four backward kernels inside the library) ----
            esize = 4 if dtype == "f32" else 2
            reps = max(1, min(args.steps, 3))
            classes = {}
            cabi.check(lib.dcnv3_set_kernel_timing(1))
            ms4 = (ctypes.c_float * 4)()
            for _ in range(reps):
                evs = []
                for l in layers:
                    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                    e0.record(stream)
                    cabi.check(lib.dcnv3_forward(*l.fwd_args, l.pref, sp))
                    e1.record(stream)
                    evs.append((("fwd_tiled",) + l.shape, e0, e1))
                torch.cuda.synchronize()
                for key, e0, e1 in evs:
                    classes.setdefault(key, []).append(e0.elapsed_time(e1))
                for l in reversed(layers):
                    cabi.check(lib.dcnv3_backward(*l.bwd_args, l.ws_bytes, l.pref, sp))
                    cabi.check(lib.dcnv3_get_kernel_timing(ms4))
                    for name, v in zip(("bwd_gather", "bwd_scatter", "bwd_redo_hot", "bwd_merge_far"), ms4):
                        if name != "bwd_merge_far" or max(l.shape[0], l.shape[1]) > 32:  # single-tile images: no merge
                            classes.setdefault((name,) + l.shape, []).append(float(v))
            cabi.check(lib.dcnv3_set_kernel_timing(0))
            table = []
            for key, ts in classes.items():
                d, h, w, c, g = key
                px = BATCH * h * w
                # algorithmic bytes of each kernel: elements per pixel x element size
                per_px = {"fwd_tiled": 2 * c + 3 * g * P, "bwd_gather": 2 * c + 6 * g * P,
                          "bwd_scatter": 2 * c + 3 * g * P}.get(d, 0)
                nbytes = px * per_px * esize
                avg = sum(ts) / len(ts)
                table.append({"kernel": f"{d} {h}x{w} C{c} G{g}", "avg_us": avg * 1e3, "launches_per_step": len(ts) // reps,
                              "algo_bytes": nbytes, "gbs": nbytes / avg * 1e-6 if avg > 0 else 0.0,
                              "share": avg * (len(ts) // reps)})
            tot = sum(t["share"] for t in table)
            for t in table:
                t["share"] = t["share"] / tot
            res["classes"] = sorted(table, key=lambda t: -t["share"])
        del layers
        torch.cuda.empty_cache()
        return res

    def measure_e2e(dtype):
        """Same step through the host-buffer C-ABI entry point: pinned host tensors in, host tensors out;
        H2D + kernels + D2H all inside the timed region."""
        tdt = torch.float32 if dtype == "f32" else torch.bfloat16
        esize = 4 if dtype == "f32" else 2
        host, h2d, d2h = [], 0, 0
        gen = torch.Generator().manual_seed(7 + rank)
        for (h, w, c, g, depth) in STAGES:
            pin = lambda *s: torch.randn(*s, generator=gen).to(tdt).pin_memory()  # noqa: E731
            x, off, go = pin(BATCH, h, w, c), pin(BATCH, h, w, g * P * 2), pin(BATCH, h, w, c)
            m = torch.softmax(torch.randn(BATCH, h, w, g, P, generator=gen), -1).reshape(BATCH, h, w, g * P).to(tdt).pin_memory()
            outs = [torch.empty_like(t).pin_memory() for t in (x, x, off, m)]
            p = cabi.make_params(x.shape, (h, w), (3, 3), (1, 1), (1, 1), (1, 1), g, GC, 1.0,
                                 cabi.F32 if dtype == "f32" else cabi.BF16)
            ptrs = [ctypes.c_void_p(t.data_ptr()) for t in (x, off, m, go, *outs)]
            host.append((ptrs, p, depth, (x, off, m, go, outs)))
            nb = (2 * x.numel() + off.numel() + m.numel()) * esize
            h2d += nb * depth
            d2h += nb * depth

        nslots = int(lib.dcnv3_host_slots())

        def step():
            i = 0
            for ptrs, p, depth, _ in host:
                for _ in range(depth):
                    # copy-in, kernels and copy-out of consecutive layers overlap on different slots
                    cabi.check(lib.dcnv3_forward_backward_host_async(*ptrs, ctypes.byref(p), local_rank, i % nslots))
                    i += 1
            cabi.check(lib.dcnv3_host_sync(local_rank))  # results are in host memory from here on

        steps = max(1, min(args.steps, args.e2e_steps))
        for _ in range(min(args.warmup, 2)):
            step()
        barrier()
        t0 = time.perf_counter()
        for _ in range(steps):
            step()
        barrier()
        dt = time.perf_counter() - t0
        if dist is not None:
            t = torch.tensor([dt], device=device)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            dt = float(t.item())
        lib.dcnv3_release_host_scratch()
        return {"value": points_per_step() * world * steps / dt, "unit": UNIT, "h2d_bytes_per_step": h2d,
                "d2h_bytes_per_step": d2h, "ms_per_step": dt / steps * 1e3, "steps": steps,
                "api": "dcnv3_forward_backward_host_async + dcnv3_host_sync (C ABI, pinned host buffers, per layer, "
                       f"{nslots} pipelined slots)"}

    sampler = ClockSampler(local_rank) if rank == 0 else None
    main = measure(args.dtype, sampler)
    other_dtype = "bf16" if args.dtype == "f32" else "f32"
    other = measure(other_dtype) if args.both_dtypes else None
    e2e = measure_e2e(args.dtype)

    if rank == 0:
        peak, peak_src = peaks()
        ms_step = main["ms_total"] / args.steps
        value = points_per_step() * world / (ms_step * 1e-3)
        top = main["classes"][0]
        traffic = None
        tpath = os.path.join(ROOT, "profiles", "traffic.json")
        if os.path.isfile(tpath):
            traffic = json.load(open(tpath)).get(f"{args.dtype}:{top['kernel']}")
        esize = 4 if args.dtype == "f32" else 2
        step_bytes = sum(sum(algo_bytes(h, w, c, g, esize)) * d for h, w, c, g, d in STAGES)
        cpu_v, cpu_ms, cpu_info = cpu_reference_run(steps=3, warmup=1, sample_batch=1, budget_s=20.0)
        cpu_info.update(value=cpu_v, unit=UNIT)
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": args.dtype, "data": "synthetic",
            "config": workload_config(args, args.dtype, world, main["launch"]),
            "clocks": sampler.summary(),
            "e2e": e2e,
            "gpu_launches": main["launches"],
            "roofline": {"bound": "hbm", "kernel": top["kernel"], "achieved": top["gbs"], "peak": peak, "unit": "GB/s",
                         "frac": top["gbs"] / peak, "traffic": traffic, "peak_source": peak_src,
                         "algo_bytes_per_launch": top["algo_bytes"], "avg_launch_us": top["avg_us"],
                         "share_of_step": top["share"]},
            "step_hbm": {"algo_bytes_per_step": step_bytes, "achieved_gbs": step_bytes / (ms_step * 1e-3) * 1e-9,
                         "frac_of_peak": step_bytes / (ms_step * 1e-3) * 1e-9 / peak},
            "kernels": [{k: (round(v, 4) if isinstance(v, float) else v) for k, v in t.items()} for t in main["classes"]],
            "cpu_baseline": cpu_info,
        }
        if other is not None:
            oms = other["ms_total"] / args.steps
            ob = sum(sum(algo_bytes(h, w, c, g, 6 - esize)) * d for h, w, c, g, d in STAGES)
            line[other_dtype] = {"value": points_per_step() * world / (oms * 1e-3), "ms_per_step": oms,
                                 "achieved_gbs": ob / (oms * 1e-3) * 1e-9, "frac_of_peak": ob / (oms * 1e-3) * 1e-9 / peak,
                                 "top_kernel": other["classes"][0]["kernel"], "top_kernel_gbs": other["classes"][0]["gbs"]}
        print(json.dumps(line))
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--dtype", default="f32", choices=["f32", "bf16"])
    ap.add_argument("--no-graph", dest="graph", action="store_false")
    ap.add_argument("--one-dtype", dest="both_dtypes", action="store_false")
    ap.add_argument("--e2e-steps", type=int, default=5)
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.gpus > 1 and "WORLD_SIZE" not in os.environ:  # convenience: self-launch one rank per GPU
        os.execvp(sys.executable, [sys.executable, "-m", "torch.distributed.run", "--nnodes=1",
                                   f"--nproc-per-node={args.gpus}", "--master-addr", "127.0.0.1",
                                   "--master-port", "29533", os.path.abspath(__file__), *sys.argv[1:]])
    if args.impl == "reference":
        run_reference(args, rank, world)
    else:
        run_ours(args, rank, world, local_rank)


if __name__ == "__main__":
    main()
