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
def cpu_reference_run(steps, warmup, batch=BATCH, budget_s=None):
    """The reference path on the host CPU: C restatement, all threads, batch `batch` of each of the 30
    layers -- every layer on its own tensors, as in the GPU arm -- forward 1..30 then backward 30..1.
    Returns (points/s, ms/step, info)."""
    import numpy as np
    from oracle import c_oracle

    c_oracle.lib(native=True)
    # every host thread this process may use (torchrun exports OMP_NUM_THREADS=1; the CPU arm ignores it)
    try:
        threads = len(os.sched_getaffinity(0))
    except AttributeError:
        threads = os.cpu_count() or 1
    rng = np.random.default_rng(0)
    layers = []
    for h, w, c, g, depth in STAGES:
        for _ in range(depth):
            x = rng.standard_normal((batch, h, w, c), dtype=np.float32)
            off = rng.standard_normal((batch, h, w, g * P * 2), dtype=np.float32)
            z = rng.standard_normal((batch, h, w, g, P), dtype=np.float32)
            np.exp(z - z.max(-1, keepdims=True), out=z)
            m = np.ascontiguousarray((z / z.sum(-1, keepdims=True)).reshape(batch, h, w, g * P))
            go = rng.standard_normal((batch, h, w, c), dtype=np.float32)
            layers.append((x, off, m, go, g))

    def one_step():
        for x, off, m, go, g in layers:
            c_oracle.forward(x, off, m, groups=g, group_channels=GC, nthreads=threads)
        for x, off, m, go, g in reversed(layers):
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
    pts = points_per_step(batch) * done
    info = {"kind": "port", "cores": threads, "batch_per_layer": batch,
            "sample": f"{done} full step(s): all 30 InternImage-T DCNv3 layers at batch {batch}, each on its own tensors, "
                      f"fwd+bwd, fp32, {points_per_step(batch)} points per step; oracle/dcnv3_oracle.c -O3 -march=native "
                      f"OpenMP, {threads} threads (TensorFlow reference not installable)"}
    return pts / dt, dt / done * 1e3, info


def run_reference(args, rank, world):
    if rank != 0:
        return
    value, ms, info = cpu_reference_run(args.steps, max(args.warmup, 1), batch=BATCH)
    info["value"] = value
    info["unit"] = UNIT
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_config(args, "f32", world), "launch": "cpu",
        "note": "the CPU arm runs ONE replica of the per-GPU workload (batch 16) on all host cores; it does not "
                "grow with --gpus",
        "cpu_baseline": info,
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


def workload_config(args, dtype, world):
    """The workload only (identical for both arms); how it is launched is reported beside it."""
    return {
        "workload": "InternImage-T DCNv3 core op fwd+bwd, all 30 layers (stages 128^2xC64/G4 x4, 64^2xC128/G8 x4, "
                    "32^2xC256/G16 x18, 16^2xC512/G32 x4), 512x512 crop",
        "batch_per_gpu": BATCH, "global_batch": BATCH * world, "kernel": "3x3 s1 d1 SAME", "group_channels": GC,
        "offset_sigma": 1.0, "parallelism": f"dp{world} (whole images sharded, no collective on the op)",
        "l2": "each layer owns its tensors (working set >> 126 MB L2); fwd 1..30 then bwd 30..1",
    }


# ------------------------------------------------------------------------------------------------
def bind_to_gpu_numa(torch, index):
    """Moves this process onto the CPUs of the NUMA node its GPU hangs off, so that the pinned host buffers
    allocated afterwards are local to the GPU's PCIe root.  Returns the node (None if unknown)."""
    try:
        pr = torch.cuda.get_device_properties(index)
        bdf = f"{pr.pci_domain_id:04x}:{pr.pci_bus_id:02x}:{pr.pci_device_id:02x}.0"
        node = int(open(f"/sys/bus/pci/devices/{bdf}/numa_node").read())
        if node < 0:
            return None
        cpus = set()
        for part in open(f"/sys/devices/system/node/node{node}/cpulist").read().strip().split(","):
            lo, _, hi = part.partition("-")
            cpus.update(range(int(lo), int(hi or lo) + 1))
        allowed = cpus & os.sched_getaffinity(0)
        if allowed:
            os.sched_setaffinity(0, allowed)
        return node
    except (OSError, ValueError, AttributeError):
        return None


SM_COUNT, SM_CLOCK_HZ = 148, 1.965e9


def by_function(classes, dtype):
    """Per kernel FUNCTION, summed over the layer shapes it is launched at in one step: time, algorithmic bytes,
    GB/s, and the on-chip ceiling that bounds it: the forward / gather kernels read 36
    corner slabs of 16 channels per (pixel, group) from shared memory at 128 B/clk/SM, the scatter kernel makes
    576 shared-memory integer atomic updates per (pixel, group) at 32 per clk per SM."""
    esize = 4 if dtype == "f32" else 2
    tpath = os.path.join(ROOT, "profiles", "traffic.json")
    traffic_tab = json.load(open(tpath)) if os.path.isfile(tpath) else {}
    agg = {}
    for t in classes:
        fn, shape = t["kernel"].split(" ", 1)
        h, w = (int(v) for v in shape.split()[0].split("x"))
        g = int(shape.split()[2][1:])
        a = agg.setdefault(fn, {"function": fn, "us_per_step": 0.0, "algo_bytes_per_step": 0, "launches_per_step": 0,
                                "ceiling_us": 0.0, "traffic": 0.0, "traffic_known": True})
        n = t["launches_per_step"]
        a["us_per_step"] += t["avg_us"] * n
        a["algo_bytes_per_step"] += t["algo_bytes"] * n
        a["launches_per_step"] += n
        pg = BATCH * h * w * g
        clk = {"fwd_tiled": 36 * 16 * esize / 128.0, "bwd_gather": 36 * 16 * esize / 128.0, "bwd_scatter": 576 / 32.0}.get(fn, 0.0)
        a["ceiling_us"] += n * pg * clk / SM_COUNT / SM_CLOCK_HZ * 1e6
        tr = traffic_tab.get(f"{dtype}:{t['kernel']}")
        if tr is None:
            a["traffic_known"] = False
        else:
            a["traffic"] += tr * n
    tot = sum(a["us_per_step"] for a in agg.values())
    out = []
    for a in sorted(agg.values(), key=lambda v: -v["us_per_step"]):
        out.append({
            "function": a["function"], "us_per_step": round(a["us_per_step"], 1), "share": round(a["us_per_step"] / tot, 4),
            "launches_per_step": a["launches_per_step"], "algo_bytes_per_step": a["algo_bytes_per_step"],
            "gbs": round(a["algo_bytes_per_step"] / a["us_per_step"] * 1e-3, 1) if a["us_per_step"] > 0 else 0.0,
            "onchip_ceiling_us": round(a["ceiling_us"], 1),
            "frac_of_onchip_ceiling": round(a["ceiling_us"] / a["us_per_step"], 4) if a["us_per_step"] > 0 else None,
            "traffic_per_launch": a["traffic"] / a["launches_per_step"] if a["traffic_known"] and a["traffic"] > 0 else None,
        })
    return out


def bf16_parity_numbers():
    """How far the two bf16 modes are from the reference's own bf16 arithmetic: max|d|/max|ref| of the forward
    over tests/golden/op_bf16_*.npz (the unmodified reference executed on bfloat16 tensors)."""
    import glob

    import numpy as np
    import torch

    import iseg_b200
    worst = {"reference_dtype_math": 0.0, "default_fp32_coordinates": 0.0}
    files = sorted(glob.glob(os.path.join(ROOT, "tests", "golden", "op_bf16_*.npz")))
    for f in files:
        z = np.load(f)
        w = lambda k: torch.from_numpy((z[k].astype(np.uint32) << 16).view(np.float32)).cuda()  # noqa: E731
        ref = w("out")
        args = [w("x").bfloat16(), w("offset").bfloat16(), w("mask").bfloat16(), [3, 3], [1, 1], "SAME", [1, 1],
                int(z["groups"]), int(z["group_channels"]), float(z["offset_scale"])]
        for key, flag in (("reference_dtype_math", True), ("default_fp32_coordinates", False)):
            out = iseg_b200.dcnv3_op(*args, reference_dtype_math=flag).float()
            worst[key] = max(worst[key], float((out - ref).abs().max() / ref.abs().max()))
    worst["fixtures"] = len(files)
    worst["note"] = ("bf16 tensors: DCNV3_FLAG_REF_DTYPE rounds every intermediate to bf16 like the reference under "
                     "mixed_bfloat16 (<= 1e-2 bar); the default bf16 mode -- the one benchmarked -- keeps coordinates and "
                     "accumulation in fp32 (forward: each finished corner weight rounded once to bf16 for its FHFMA product) and is "
                     "held to 1e-2 against the fp32 oracle on the bf16-rounded inputs instead")
    return worst


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
                        # whole-image scatter tiles (<= 32x32): no merge launch, and no redo launch in fp32
                        if (name in ("bwd_gather", "bwd_scatter") or max(l.shape[0], l.shape[1]) > 32
                                or (name == "bwd_redo_hot" and dtype == "bf16")):
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

    def pcie_ceiling():
        """What the host lets every rank copy at the same time: all ranks run one pinned 256 MB H2D and one D2H
        copy concurrently (duplex, separate streams); GB/s per direction, the slowest rank's."""
        nb = 256 << 20
        hin, hout = torch.empty(nb, dtype=torch.uint8).pin_memory(), torch.empty(nb, dtype=torch.uint8).pin_memory()
        din, dout = torch.empty(nb, dtype=torch.uint8, device=device), torch.empty(nb, dtype=torch.uint8, device=device)
        s1, s2 = torch.cuda.Stream(device=device), torch.cuda.Stream(device=device)
        best = 0.0
        for _ in range(4):
            barrier()
            t0 = time.perf_counter()
            with torch.cuda.stream(s1):
                din.copy_(hin, non_blocking=True)
            with torch.cuda.stream(s2):
                hout.copy_(dout, non_blocking=True)
            s1.synchronize(), s2.synchronize()
            dt = time.perf_counter() - t0
            if dist is not None:
                t = torch.tensor([dt], device=device)
                dist.all_reduce(t, op=dist.ReduceOp.MAX)
                dt = float(t.item())
            best = max(best, nb / dt * 1e-9)
        return best

    def measure_e2e(dtype):
        """Same step through the host-buffer C-ABI entry point: pinned host tensors in, host tensors out;
        H2D + kernels + D2H all inside the timed region."""
        numa = bind_to_gpu_numa(torch, local_rank) if world > 1 else None  # pinned pages first-touch on the GPU's node
        ceiling = pcie_ceiling()
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
                "pcie_gbs_per_direction_all_ranks_concurrent": ceiling,
                "achieved_gbs_per_direction": h2d / (dt / steps) * 1e-9,
                "frac_of_pcie_ceiling": h2d / (dt / steps) * 1e-9 / ceiling if ceiling > 0 else None,
                "numa_node": numa,
                "api": "dcnv3_forward_backward_host_async + dcnv3_host_sync (C ABI, pinned host buffers, per layer, "
                       f"{nslots} pipelined slots)"}

    def measure_gather(dtype):
        """NCCL all_gather_into_tensor of every rank's stage-1 output shard (what the reference's
        experimental_local_results + concat does, core_predict.py:136-153): alone, and issued on a side stream
        while the next layer's forward runs on the compute stream."""
        tdt = torch.float32 if dtype == "f32" else torch.bfloat16
        h, w, c, g, _ = STAGES[0]
        layer = Layer(torch, cabi, h, w, c, g, dtype, 999 + rank, device)
        full = torch.empty((world * BATCH, h, w, c), dtype=tdt, device=device)
        comm, comp = torch.cuda.Stream(device=device), torch.cuda.Stream(device=device)
        sp = ctypes.c_void_p(comp.cuda_stream)
        nbytes = layer.out.numel() * layer.out.element_size()

        def fwd():
            with torch.cuda.stream(comp):
                cabi.check(lib.dcnv3_forward(*layer.fwd_args, layer.pref, sp))

        def gather_only():
            with torch.cuda.stream(comm):
                dist.all_gather_into_tensor(full, layer.out)

        def both():
            fwd()
            gather_only()

        def timeit(fn, reps=20):
            for _ in range(3):
                fn()
            barrier()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            for _ in range(reps):
                fn()
                torch.cuda.current_stream().wait_stream(comm)
                torch.cuda.current_stream().wait_stream(comp)
            b.record()
            barrier()
            t = torch.tensor([a.elapsed_time(b) / reps], device=device)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            return float(t.item())

        t_f, t_g, t_b = timeit(fwd), timeit(gather_only), timeit(both)
        ok = bool(torch.equal(full[rank * BATCH:(rank + 1) * BATCH], layer.out))
        recv = (world - 1) * nbytes
        return {"what": "all_gather_into_tensor of the stage-1 output shards (NCCL over NVLink), max over ranks",
                "shard_bytes": nbytes, "gathered_bytes": world * nbytes, "ms_gather": t_g, "ms_forward": t_f,
                "ms_forward_and_gather_overlapped": t_b, "recv_gbs_per_gpu": recv / (t_g * 1e-3) * 1e-9,
                "nvlink_nominal_gbs_per_direction": 900.0, "own_shard_bit_exact": ok}

    sampler = ClockSampler(local_rank) if rank == 0 else None
    main = measure(args.dtype, sampler)
    other_dtype = "bf16" if args.dtype == "f32" else "f32"
    other = measure(other_dtype) if args.both_dtypes else None
    if args.kernels_only:  # developer A/B aid (tools/ab_step.sh): the device-timed step and its functions, nothing else
        if rank == 0:
            out = {}
            for dt, m in ((args.dtype, main), (other_dtype, other)):
                if m is not None:
                    out[dt] = {"ms_per_step": round(m["ms_total"] / args.steps, 4),
                               **{f["function"]: round(f["us_per_step"], 1) for f in by_function(m["classes"], dt)}}
            print(json.dumps(out))
        return
    e2e = measure_e2e(args.dtype)

    gather = measure_gather(args.dtype) if dist is not None else None
    parity = bf16_parity_numbers() if rank == 0 else None

    if rank == 0:
        peak, peak_src = peaks()
        ms_step = main["ms_total"] / args.steps
        value = points_per_step() * world / (ms_step * 1e-3)
        esize = 4 if args.dtype == "f32" else 2
        step_bytes = sum(sum(algo_bytes(h, w, c, g, esize)) * d for h, w, c, g, d in STAGES)
        functions = by_function(main["classes"], args.dtype)
        top = functions[0]
        if world == 1:
            cpu_v, cpu_ms, cpu_info = cpu_reference_run(steps=2, warmup=1, batch=BATCH, budget_s=25.0)
            cpu_info.update(value=cpu_v, unit=UNIT, ms_per_step=cpu_ms)
        else:  # measured on rank 0 at N = 1 only: the other ranks' host threads would share the cores
            cpu_info = {"value": None, "unit": UNIT, "kind": "port", "cores": 0, "sample": "not measured at N > 1"}
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": args.dtype, "data": "synthetic",
            "config": workload_config(args, args.dtype, world), "launch": main["launch"],
            "clocks": sampler.summary(),
            "e2e": e2e,
            "gpu_launches": main["launches"],
            # the dominant kernel FUNCTION of the step, summed over the four layer shapes it runs at
            "roofline": {"bound": "hbm", "kernel": top["function"], "achieved": top["gbs"], "peak": peak, "unit": "GB/s",
                         "frac": top["gbs"] / peak, "traffic": top["traffic_per_launch"], "peak_source": peak_src,
                         "launches_per_step": top["launches_per_step"],
                         "algo_bytes_per_launch": top["algo_bytes_per_step"] / top["launches_per_step"],
                         "avg_launch_us": top["us_per_step"] / top["launches_per_step"],
                         "share_of_step": top["share"], "onchip_ceiling_frac": top["frac_of_onchip_ceiling"],
                         "traffic_note": "ncu --set full dram__bytes_read+write per launch, launch-weighted mean over the layer "
                                         "shapes (profiles/traffic.json; null if a shape was not captured).  It can sit "
                                         "below the algorithmic bytes: a kernel's last writes are still dirty in the "
                                         "126 MB L2 when it ends and inputs written by the previous kernel are read from L2"},
            "step_hbm": {"algo_bytes_per_step": step_bytes, "achieved_gbs": step_bytes / (ms_step * 1e-3) * 1e-9,
                         "frac_of_peak": step_bytes / (ms_step * 1e-3) * 1e-9 / peak},
            "functions": functions,
            "kernels": [{k: (round(v, 4) if isinstance(v, float) else v) for k, v in t.items()} for t in main["classes"]],
            "bf16_parity": parity,
            "cpu_baseline": cpu_info,
        }
        if gather is not None:
            line["gather"] = gather
        if other is not None:
            oms = other["ms_total"] / args.steps
            ob = sum(sum(algo_bytes(h, w, c, g, 6 - esize)) * d for h, w, c, g, d in STAGES)
            of = by_function(other["classes"], other_dtype)
            line[other_dtype] = {"value": points_per_step() * world / (oms * 1e-3), "ms_per_step": oms,
                                 "achieved_gbs": ob / (oms * 1e-3) * 1e-9, "frac_of_peak": ob / (oms * 1e-3) * 1e-9 / peak,
                                 "functions": of}
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
    ap.add_argument("--kernels-only", action="store_true", help="developer aid: device-timed step only, short output")
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
