#!/usr/bin/env python
"""bench.py -- throughput of the PWC-Net warp + cost-volume hot path on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl native|reference] [--workload ...]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

Workload (BASELINE.json configs[1], the configuration the metric is quoted on that fits one GPU):
the fused warp+correlation microbenchmark at the level-2 shape (B=32, C=32, 96x112, md=4) and the
level-6 shape (B=32, C=196, 6x7) of a batch of 32 synthetic 448x384 image pairs, forward +
backward.  One step = both shapes, fwd + bwd, for the 32 pairs of one GPU; `value` = pairs/s over
all GPUs (weak scaling: 32 pairs per GPU, sharded by image pair, no data-path collective).

Arms:
  --impl native     the CUDA path through the public Python API / C ABI (libpwc_b200.so).
  --impl reference  the reference's PyTorch-level path (modules.WarpingLayer + CostVolumeLayer,
                    modules.py:25-74) restated in oracle/torch_ref.py, on the host CPU cores with
                    autograd, on a bounded sample of the same workload (rank 0 only).

The single JSON line carries `roofline` (fused forward kernel at the level-2 shape, algorithmic
bytes / CUDA-event time vs MEASURED_PEAKS.json), `e2e` (same step with every input copied from
pinned host memory and every result copied back inside the timed region), `cpu_baseline`,
`clocks` and `gpu_launches`.
"""
import argparse
import json
import os
import statistics
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

CANON = dict(pad_size=4, kernel_size=1, max_displacement=4, stride1=1, stride2=1)
SHAPES = {  # name: (B, C, H, W)
    "level2": (32, 32, 96, 112),
    "level6": (32, 196, 6, 7),
}
# the five correlation calls of one 384x448 pair
PYRAMID_384x448 = [(192, 6, 7), (128, 12, 14), (96, 24, 28), (64, 48, 56), (32, 96, 112)]
PAIRS_PER_GPU = 32
CPU_SAMPLE_PAIRS = 8


def fwd_bytes(B, C, H, W, D2=81):
    """Algorithmic HBM bytes of one fused forward: read f1, f2, flow once, write D*D outputs
    (4*(2C + 2 + 81) per pixel)."""
    return 4 * (2 * C + 2 + D2) * B * H * W


def bwd_bytes(B, C, H, W, D2=81):
    """read grad_out, f1, f2, flow; write g1, g2, gflow (4*(81 + 4C + 4) per pixel)."""
    return 4 * (D2 + 4 * C + 4) * B * H * W


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return json.load(open(p)), "measured (MEASURED_PEAKS.json)"
        except Exception:
            pass
    return {"hbm_gbs": 6650.0}, "fallback (B200_PROFILING.md: 6.65 TB/s)"


# ------------------------------------------------------------------------------------------------
# clocks sampler (pynvml, falls back to nvidia-smi)
# ------------------------------------------------------------------------------------------------
class ClockSampler:
    REASONS = {
        0x1: "gpu_idle", 0x2: "applications_clocks_setting", 0x4: "sw_power_cap", 0x8: "hw_slowdown",
        0x10: "sync_boost", 0x20: "sw_thermal_slowdown", 0x40: "hw_thermal_slowdown",
        0x80: "hw_power_brake_slowdown", 0x100: "display_clock_setting",
    }

    def __init__(self, index, period_s=0.002):
        self.index, self.period = index, period_s
        self.samples, self.reasons = [], set()
        self.max_mhz = None
        self._stop = threading.Event()
        self._thr = None
        self._nvml = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self._nvml = pynvml
            self._h = pynvml.nvmlDeviceGetHandleByIndex(self._physical_index(index))
            self.max_mhz = float(pynvml.nvmlDeviceGetMaxClockInfo(self._h, pynvml.NVML_CLOCK_SM))
        except Exception:
            self._nvml = None

    @staticmethod
    def _physical_index(i):
        vis = os.environ.get("CUDA_VISIBLE_DEVICES")
        if vis:
            try:
                return int(vis.split(",")[i])
            except Exception:
                return i
        return i

    def _loop(self):
        nv = self._nvml
        while not self._stop.is_set():
            try:
                mhz = float(nv.nvmlDeviceGetClockInfo(self._h, nv.NVML_CLOCK_SM))
                try:
                    mask = int(nv.nvmlDeviceGetCurrentClocksEventReasons(self._h))
                except Exception:
                    mask = int(nv.nvmlDeviceGetCurrentClocksThrottleReasons(self._h))
                self.samples.append(mhz)
                for bit, name in self.REASONS.items():
                    if mask & bit and name != "gpu_idle":
                        self.reasons.add(name)
            except Exception:
                pass
            self._stop.wait(self.period)

    def start(self):
        if self._nvml is not None:
            self._thr = threading.Thread(target=self._loop, daemon=True)
            self._thr.start()
        return self

    def stop(self):
        self._stop.set()
        if self._thr is not None:
            self._thr.join(timeout=2.0)
        if not self.samples:
            self._nvidia_smi_once()
        return {
            "sm_mhz": statistics.median(self.samples) if self.samples else None,
            "sm_max_mhz": self.max_mhz,
            "reasons": sorted(self.reasons),
            "samples": len(self.samples),
        }

    def _nvidia_smi_once(self):
        try:
            import subprocess
            out = subprocess.run(["nvidia-smi", f"--id={self._physical_index(self.index)}",
                                  "--query-gpu=clocks.sm,clocks.max.sm", "--format=csv,noheader,nounits"],
                                 capture_output=True, text=True, timeout=10).stdout.strip().split(",")
            self.samples.append(float(out[0]))
            self.max_mhz = float(out[1])
        except Exception:
            pass


# ------------------------------------------------------------------------------------------------
# reference arm / cpu baseline: the reference's PyTorch-level path on the host cores
# ------------------------------------------------------------------------------------------------
_REF_DIR = os.path.join(ROOT, "baseline", "_ref")


def reference_modules():
    """The reference's own, unmodified modules.py / utils.py (WarpingLayer, CostVolumeLayer), installed
    into the git-ignored baseline/_ref/ by __graft_entry__.build() when /root/reference is present (the
    reference has no setup.py, so `pip install --target baseline/_ref` does not apply; the two files are
    copied as they are).  Returns (warp, cost_volume) callables or None.  The only intervention is the one
    the parity pins use as well: grid_sample is called with align_corners=True, the torch-0.4.0 behaviour
    the reference was written against."""
    if not (os.path.exists(os.path.join(_REF_DIR, "modules.py")) and os.path.exists(os.path.join(_REF_DIR, "utils.py"))):
        return None
    try:
        import importlib.util
        import types

        import torch
        import torch.nn.functional as F
        mods = {}
        for name in ("utils", "modules"):       # loaded under private names: nothing else sees them
            spec = importlib.util.spec_from_file_location(f"_pwc_reference_{name}", os.path.join(_REF_DIR, f"{name}.py"))
            m = importlib.util.module_from_spec(spec)
            if name == "modules":
                sys.modules["utils"] = mods["utils"]        # modules.py:8 `from utils import get_grid`
            try:
                spec.loader.exec_module(m)
            finally:
                if name == "modules":
                    sys.modules.pop("utils", None)
            mods[name] = m
        ref = mods["modules"]
        fns = types.SimpleNamespace(**{k: getattr(F, k) for k in dir(F) if not k.startswith("__")})
        fns.grid_sample = lambda x, grid: F.grid_sample(x, grid, mode="bilinear", padding_mode="zeros",
                                                        align_corners=True)
        ref.F = fns
        a = types.SimpleNamespace(device=torch.device("cpu"), search_range=4)
        return ref.WarpingLayer(a), ref.CostVolumeLayer(a)
    except Exception:
        return None


def cpu_path_step(inputs, backward=True, ref=None):
    """WarpingLayer -> CostVolumeLayer (modules.py:31-42, :52-74) fwd (+ autograd bwd) on CPU: the
    reference's own modules when available (`ref`), else their port in oracle/torch_ref.py."""
    if ref is None:
        from oracle import torch_ref as tr
        warp, cost = tr.warping_layer_port, (lambda a, b: tr.cost_volume_layer_port(a, b, 4))
    else:
        warp, cost = ref
    pairs = 0
    for (f1, f2, flow, gout) in inputs:
        if backward:
            f1 = f1.detach().requires_grad_()
            f2 = f2.detach().requires_grad_()
            flow = flow.detach().requires_grad_()
        out = cost(f1, warp(f2, flow))
        if backward:
            out.backward(gout)
        pairs = max(pairs, f1.shape[0])
    return pairs


def make_cpu_inputs(torch, sample_pairs, seed=0):
    g = torch.Generator().manual_seed(seed)
    inputs = []
    for name, (B, C, H, W) in SHAPES.items():
        b = min(B, sample_pairs)
        f1 = torch.randn(b, C, H, W, generator=g)
        f2 = torch.randn(b, C, H, W, generator=g)
        flow = 2.0 * torch.randn(b, 2, H, W, generator=g)
        gout = torch.randn(b, 81, H, W, generator=g)
        inputs.append((f1, f2, flow, gout))
    return inputs


def time_cpu_path(steps, warmup, sample_pairs):
    import torch
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    inputs = make_cpu_inputs(torch, sample_pairs)
    ref = reference_modules()
    for _ in range(warmup):
        cpu_path_step(inputs, ref=ref)
    t0 = time.perf_counter()
    for _ in range(steps):
        pairs = cpu_path_step(inputs, ref=ref)
    dt = time.perf_counter() - t0
    what = ("the reference's own modules.WarpingLayer + CostVolumeLayer (unmodified files in baseline/_ref, "
            "grid_sample at align_corners=True)" if ref is not None else
            "oracle/torch_ref.py port of modules.WarpingLayer+CostVolumeLayer")
    return {
        "value": pairs * steps / dt, "unit": "pairs/s", "cores": cores, "kind": "reference" if ref is not None else "port",
        "sample": (f"{sample_pairs} of {PAIRS_PER_GPU} pairs per step, both shapes, fwd+bwd (autograd), "
                   f"{steps} steps after {warmup} warm-up; {what}, torch CPU"),
        "ms_per_step": 1e3 * dt / steps,
    }


def cpu_model_name():
    try:
        for ln in open("/proc/cpuinfo"):
            if ln.startswith("model name"):
                return ln.split(":", 1)[1].strip()
    except Exception:
        pass
    return "unknown"


def base_config():
    return {
        "workload": ("cfg2 micro: fused warp+corr fwd+bwd at level-2 (B=32,C=32,96x112) and level-6 "
                     "(B=32,C=196,6x7) shapes of 32 synthetic 448x384 pairs per GPU, md=4 (81 displacements)"),
        "pairs_per_gpu": PAIRS_PER_GPU,
        "corr": CANON,
        "flow": "iid N(0, 2^2) px (worst-case gather)",
        "parallelism": "data-parallel by image pair, no collective",
    }


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    # same step / warm-up counts as the native arm; each step is a bounded sample (8 of the 32 pairs,
    # ~0.2 s of host work), the step count is capped so that the run ends within a few minutes
    steps = max(1, min(args.steps, 200))
    warmup = max(0, min(args.warmup, 10))
    res = time_cpu_path(steps, warmup, CPU_SAMPLE_PAIRS)
    cfg = base_config()
    cfg["cpu"] = cpu_model_name()
    line = {
        "impl": "reference", "metric": "image_pairs_per_sec", "value": res["value"], "unit": "pairs/s",
        "n_gpus": args.gpus, "steps": steps, "warmup": warmup, "ms_per_step": res["ms_per_step"],
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
        "data": "synthetic", "config": cfg,
        "cpu_baseline": {k: res[k] for k in ("value", "unit", "cores", "kind", "sample")},
        "e2e": {"value": res["value"], "unit": "pairs/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    emit(line)
    return 0


# ------------------------------------------------------------------------------------------------
# native arm
# ------------------------------------------------------------------------------------------------
def run_native(args):
    import torch
    import torch.distributed as dist

    import pwc_net_pytorch_b200 as pkg
    from pwc_net_pytorch_b200 import _lib
    from pwc_net_pytorch_b200 import parallel

    if not torch.cuda.is_available():
        raise SystemExit("bench.py --impl native needs a CUDA device (no CPU fallback exists)")
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        # NCCL writes its banner / debug lines to stdout by default; stdout carries exactly one JSON line
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
        dist.init_process_group("nccl", device_id=dev)
    lib = _lib.load()
    torch.manual_seed(rank)

    op = pkg.FusedWarpCorrelation(**CANON)
    NSETS = 3   # rotate input sets so consecutive steps never reuse the same buffers from L2

    def make_set(flow_kind="iid"):
        s = {}
        for name, (B, C, H, W) in SHAPES.items():
            f1 = torch.randn(B, C, H, W, device=dev)
            f2 = torch.randn(B, C, H, W, device=dev)
            if flow_kind == "iid":
                flow = 2.0 * torch.randn(B, 2, H, W, device=dev)
            else:
                coarse = 2.0 * torch.randn(B, 2, max(2, H // 8 + 1), max(2, W // 8 + 1), device=dev)
                flow = torch.nn.functional.interpolate(coarse, size=(H, W), mode="bilinear", align_corners=True)
            gout = torch.randn(B, 81, H, W, device=dev)
            s[name] = tuple(t.contiguous() for t in (f1, f2, flow, gout))
        return s

    sets = [make_set() for _ in range(NSETS)]
    ev_pairs = []

    # The two shapes of the workload are independent.  The level-2 kernels are persistent (one CTA per SM,
    # 9.08 tiles per CTA: 12 of the 148 CTAs carry a tenth tile), so every one of them ends with a tail in
    # which most SMs idle; the coarse level runs on a second stream and fills those tails.
    side = torch.cuda.Stream()

    def one_level(s, name, record, ext=None):
        f1, f2, flow, gout = s[name]
        a = f1.requires_grad_()
        b = f2.requires_grad_()
        f = flow.requires_grad_()
        a.grad = b.grad = f.grad = None
        if name == "level2" and (record or ext is not None):
            # events around the roofline kernel: plain events in eager mode; under graph capture
            # "external" events (cudaEventRecordExternal), which become event-record nodes of the graph and
            # can be read after a replay
            e0, e1 = ext if ext is not None else (torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True))
            e0.record()
            out = op(a, b, f)
            e1.record()
            if ext is None:
                ev_pairs.append((e0, e1))
        else:
            out = op(a, b, f)
        out.backward(gout)

    def step(i, record=False, ext=None):
        s = sets[i % NSETS]
        main_s = torch.cuda.current_stream()
        side.wait_stream(main_s)                 # fork: the step starts when the previous one has finished
        one_level(s, "level2", record, ext)      # launched first: owns the SMs
        with torch.cuda.stream(side):
            one_level(s, "level6", record)       # scheduled wherever the level-2 kernels leave SMs free
        main_s.wait_stream(side)                 # join: the step ends when both shapes are done

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- warm-up (also counts launches per step) ----
    warmup = max(3, args.warmup)
    step(0)
    torch.cuda.synchronize()
    l0 = lib.pwc_launch_count()
    step(1)
    torch.cuda.synchronize()
    launches_per_step = int(lib.pwc_launch_count() - l0)
    step(2)
    torch.cuda.synchronize()

    # ---- capture one CUDA graph per input set (the library's entry points are capturable: no
    # allocation, no sync); replay removes the Python/ctypes launch overhead from the timed region ----
    graphs = None
    graph_events = []
    NGRAPH = 2 * NSETS      # two graphs per input set: six in-region samples of the roofline kernel
    if not args.eager:
        try:
            graphs = []
            cap_stream = torch.cuda.Stream()
            cap_stream.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(cap_stream):
                for i in range(NSETS):
                    step(i)            # warm the capture stream's allocator pool
            torch.cuda.current_stream().wait_stream(cap_stream)
            torch.cuda.synchronize()
            for i in range(NGRAPH):
                ext = (torch.cuda.Event(enable_timing=True, external=True),
                       torch.cuda.Event(enable_timing=True, external=True))
                g = torch.cuda.CUDAGraph()
                with torch.cuda.graph(g):
                    step(i, ext=ext)
                graphs.append(g)
                graph_events.append(ext)
            for g in graphs:
                g.replay()
            torch.cuda.synchronize()
        except Exception as e:      # eager fallback keeps the measurement valid, just launch-bound
            graphs = None
            graph_error = repr(e)
            torch.cuda.synchronize()

    def run_step(i, record=False):
        if graphs is None:
            step(i, record=record)
        else:
            graphs[i % NGRAPH].replay()

    for i in range(3, warmup + 3):
        run_step(i)

    # ---- graph replay: the roofline kernel is also timed alone (reported as avg_launch_ms_alone), back
    # to back behind a device-side sleep so that no CPU launch gap is included; the figure used for the
    # roofline comes from the event-record nodes inside the replayed graphs (see below) ----
    fwd_ms = []
    fwd_alone_ms = None
    if graphs is not None:
        pairs_ev = []
        torch.cuda._sleep(20_000_000)
        with torch.no_grad():
            for i in range(30):
                f1, f2, flow, _ = sets[i % NSETS]["level2"]
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                op(f1, f2, flow)
                e1.record()
                pairs_ev.append((e0, e1))
        torch.cuda.synchronize()
        fwd_ms = [a.elapsed_time(b) for a, b in pairs_ev]
        fwd_alone_ms = sum(fwd_ms) / len(fwd_ms)
    # The same kernel with the flow SURVEY.md 8(d) calls typical -- 8x-downsampled noise, bilinearly upsampled, which is
    # what model.py:78 hands to every level -- timed alone the same way; the headline roofline stays on the i.i.d. flow.
    fwd_smooth_ms = None
    try:
        if graphs is None:
            raise RuntimeError("eager mode: keep the launch list of the timed step clean")
        torch.manual_seed(1234 + rank)
        smooth_sets = [make_set("smooth")["level2"] for _ in range(NSETS)]
        torch.cuda._sleep(20_000_000)
        pairs_sm = []
        with torch.no_grad():
            for i in range(33):
                f1, f2, flow, _ = smooth_sets[i % NSETS]
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                op(f1, f2, flow)
                e1.record()
                pairs_sm.append((e0, e1))
        torch.cuda.synchronize()
        sm = [a.elapsed_time(b) for a, b in pairs_sm[3:]]
        fwd_smooth_ms = sum(sm) / len(sm)
        del smooth_sets
    except Exception:
        fwd_smooth_ms = None

    # ---- timed region: device-resident inputs ----
    barrier()
    sampler = ClockSampler(local).start()
    t_start, t_end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t_start.record()
    for i in range(args.steps):
        run_step(i, record=True)
    t_end.record()
    barrier()
    clocks = sampler.stop()
    ms_total = t_start.elapsed_time(t_end)
    roofline_timing = "alone, right before the timed region"
    if graphs is None:      # events recorded around the fused forward launch inside the timed region
        fwd_ms = [a.elapsed_time(b) for a, b in ev_pairs]
        roofline_timing = f"CUDA events inside the timed region, every step ({len(fwd_ms)} samples)"
    else:
        try:                # event-record nodes of the graphs: the last replay of each graph = the last steps of the region
            used = graph_events[:min(NGRAPH, args.steps)]
            fwd_ms = [a.elapsed_time(b) for a, b in used]
            roofline_timing = (f"CUDA event-record nodes inside the replayed graphs, last {len(fwd_ms)} steps of the "
                               "timed region")
        except Exception:
            pass
    fwd_l2_ms = sum(fwd_ms) / len(fwd_ms)

    # whole-job pairs/s = pairs of all ranks / slowest rank's device time (no data-path collective)
    value, ms_total_max = parallel.job_throughput(PAIRS_PER_GPU * args.steps, ms_total, dev)

    # ---- e2e: same step, every input from pinned host memory, every result back to the host ----
    host = {}
    h2d = d2h = 0
    for name, (B, C, H, W) in SHAPES.items():
        hin = [torch.randn(B, C, H, W).pin_memory(), torch.randn(B, C, H, W).pin_memory(),
               (2.0 * torch.randn(B, 2, H, W)).pin_memory(), torch.randn(B, 81, H, W).pin_memory()]
        hout = [torch.empty(B, 81, H, W).pin_memory(), torch.empty(B, C, H, W).pin_memory(),
                torch.empty(B, C, H, W).pin_memory(), torch.empty(B, 2, H, W).pin_memory()]
        host[name] = (hin, hout)
        h2d += sum(x.numel() * 4 for x in hin)
        d2h += sum(x.numel() * 4 for x in hout)

    # Triple-buffered pipeline: H2D of step i+1, compute of step i and D2H of step i-1 run on three
    # streams; every step's copies are inside the timed region.
    main = torch.cuda.current_stream()
    h2d_s, d2h_s = torch.cuda.Stream(), torch.cuda.Stream()
    NBUF = 3
    dev_in = [{n: [torch.empty_like(x, device=dev) for x in host[n][0]] for n in SHAPES} for _ in range(NBUF)]
    for b in range(NBUF):
        for n in SHAPES:
            for t in dev_in[b][n][:3]:
                t.requires_grad_()
    ev_in = [torch.cuda.Event() for _ in range(NBUF)]
    ev_done = [torch.cuda.Event() for _ in range(NBUF)]
    for e in ev_done:
        e.record(main)

    def e2e_compute(b):
        results = {}
        for n in ("level2", "level6"):
            f1, f2, flow, gout = dev_in[b][n]
            f1.grad = f2.grad = flow.grad = None
            out = op(f1, f2, flow)
            out.backward(gout)
            results[n] = [out.detach(), f1.grad, f2.grad, flow.grad]
        return results

    # The compute part of the end-to-end step is replayed from one CUDA graph per buffer set as well (the public API is
    # capturable), so that eight ranks launching from Python on one host do not compete for its cores inside the timed
    # region; `--eager` keeps the Python launches.  The copies stay outside the graphs: they are the thing measured.
    e2e_graphs, e2e_static = None, None
    ev_d2h = [torch.cuda.Event() for _ in range(NBUF)]
    for e in ev_d2h:
        e.record(main)
    if graphs is not None:
        try:
            cap = torch.cuda.Stream()
            cap.wait_stream(main)
            with torch.cuda.stream(cap):
                for b in range(NBUF):
                    e2e_compute(b)
            main.wait_stream(cap)
            torch.cuda.synchronize()
            e2e_graphs, e2e_static = [], []
            for b in range(NBUF):
                g = torch.cuda.CUDAGraph()
                with torch.cuda.graph(g):
                    res = e2e_compute(b)
                e2e_graphs.append(g)
                e2e_static.append(res)
            torch.cuda.synchronize()
        except Exception:
            e2e_graphs, e2e_static = None, None
            torch.cuda.synchronize()

    def e2e_step(i):
        b = i % NBUF
        with torch.cuda.stream(h2d_s):
            h2d_s.wait_event(ev_done[b])          # buffer b was last used by step i - NBUF
            with torch.no_grad():
                for n in SHAPES:
                    for dst, src in zip(dev_in[b][n], host[n][0]):
                        dst.copy_(src, non_blocking=True)
            ev_in[b].record(h2d_s)
        main.wait_event(ev_in[b])
        if e2e_graphs is not None:
            main.wait_event(ev_d2h[b])            # the graph's static results of step i - NBUF have been copied out
            e2e_graphs[b].replay()
            results = e2e_static[b]
        else:
            results = e2e_compute(b)
        ev_done[b].record(main)
        with torch.cuda.stream(d2h_s):
            d2h_s.wait_event(ev_done[b])
            for n in SHAPES:
                for r, hdst in zip(results[n], host[n][1]):
                    hdst.copy_(r, non_blocking=True)
                    if e2e_graphs is None:
                        r.record_stream(d2h_s)
            ev_d2h[b].record(d2h_s)

    e2e_steps = max(3, min(args.steps, 40))
    for i in range(4):
        e2e_step(i)
    main.wait_stream(d2h_s)
    barrier()
    e_start, e_end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e_start.record()
    for i in range(e2e_steps):
        e2e_step(i)
    main.wait_stream(d2h_s)
    main.wait_stream(h2d_s)
    e_end.record()
    barrier()
    e2e_value, _ = parallel.job_throughput(PAIRS_PER_GPU * e2e_steps, e_start.elapsed_time(e_end), dev)
    e2e_mode = "cuda_graph_replay" if e2e_graphs is not None else "eager"

    # ---- whole-network legs (all ranks) + per-kernel breakdown (rank 0), outside the timed region ----
    extras = {}
    legs = {}
    if not args.no_extras:
        e2e_graphs = e2e_static = None
        dev_in.clear()       # release the e2e buffers before the network legs
        host.clear()
        torch.cuda.empty_cache()
        legs = measure_network_legs(torch, dist, dev, rank, world, args.dump_kernels)
    if rank == 0 and not args.no_extras:
        extras = measure_extras(torch, pkg, dev, sets, make_set)
        extras.update(legs)

    if rank == 0:
        peaks, peak_src = measured_peaks()
        peak = float(peaks.get("hbm_gbs", 6650.0))
        B, C, H, W = SHAPES["level2"]
        ach = fwd_bytes(B, C, H, W) / (fwd_l2_ms * 1e-3) / 1e9
        traffic = None
        tinfo = {}
        tpath = os.path.join(ROOT, "profiles", "traffic.json")
        if os.path.exists(tpath):
            try:
                tinfo = json.load(open(tpath))
                traffic = tinfo.get("warpcorr_fwd_level2_dram_bytes")
            except Exception:
                traffic, tinfo = None, {}
        # the compute floor beside the HBM one: 81 * C FMAs per pixel on the fp32 pipe at its
        # 128 lane-FMA/clk/SM peak, and at the ~57 % a register-tiled band correlation sustains (measured rates in
        # profiles/r02_forward_limits.txt); at the level-2 shape the latter is 42 us, i.e. 0.74 of the HBM roofline
        fma_floor = None
        if clocks.get("sm_max_mhz"):
            sms_ = torch.cuda.get_device_properties(dev).multi_processor_count
            B2, C2, H2, W2 = SHAPES["level2"]
            t_peak = 81.0 * C2 * B2 * H2 * W2 / (sms_ * 128 * float(clocks["sm_max_mhz"]) * 1e6) * 1e3
            fma_floor = {"ms_at_fma_peak": t_peak, "ms_at_sustained_simt_rate": t_peak / 0.57,
                         "frac_of_hbm_roofline_reachable": fwd_bytes(*SHAPES["level2"]) / (t_peak / 0.57) / 1e6 / peak,
                         "source": "profiles/r02_forward_limits.txt"}
        # what actually binds the kernel: the shared-memory data pipe, one 128-byte
        # wavefront per SM and clock; wavefront count from the same ncu capture as `traffic`
        lsu = None
        if tinfo.get("warpcorr_fwd_level2_lsu_wavefronts") and clocks.get("sm_max_mhz"):
            wf = float(tinfo["warpcorr_fwd_level2_lsu_wavefronts"])
            sms = torch.cuda.get_device_properties(dev).multi_processor_count
            peak_wf = sms * float(clocks["sm_max_mhz"]) * 1e6
            lsu = {"resource": "shared-memory data pipe (l1tex__data_pipe_lsu_wavefronts, 1 wavefront = 128 B per SM and clock)",
                   "wavefronts_per_launch": wf, "peak_wavefronts_per_s": peak_wf,
                   "min_ms_at_peak": wf / peak_wf * 1e3, "frac": (wf / peak_wf * 1e3) / fwd_l2_ms,
                   "source": tinfo.get("source")}
        cpu = None
        if not args.no_cpu_baseline:
            cpu = time_cpu_path(5, 1, CPU_SAMPLE_PAIRS)
            cpu["cpu"] = cpu_model_name()
        cfg = base_config()
        cfg["l2_policy"] = (f"inputs rotated over {NSETS} buffer sets (0.27 GB) and ~0.5 GB touched per "
                            "step, larger than the 126 MB L2")
        line = {
            "metric": "image_pairs_per_sec", "value": value, "unit": "pairs/s", "n_gpus": world,
            "steps": args.steps, "warmup": warmup, "ms_per_step": ms_total_max / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
            "data": "synthetic", "config": cfg, "clocks": clocks,
            "e2e": {"value": e2e_value, "unit": "pairs/s", "h2d_bytes_per_step": h2d,
                    "d2h_bytes_per_step": d2h, "steps": e2e_steps,
                    "launch_mode": e2e_mode},
            "gpu_launches": launches_per_step * args.steps,
            "launch_mode": "cuda_graph_replay" if graphs is not None else "eager",
            "roofline": {
                "kernel": "warpcorr_fwd_tma_kernel (fused warp+corr forward, level-2 shape B=32 C=32 96x112)",
                "bound": "hbm", "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak,
                "traffic": traffic, "traffic_source": tinfo.get("source"),
                "traffic_read_write": [tinfo.get("warpcorr_fwd_level2_dram_read_bytes"),
                                       tinfo.get("warpcorr_fwd_level2_dram_write_bytes")],
                "binding_resource": lsu, "compute_floor": fma_floor, "peak_source": peak_src,
                "typical_flow": (None if not fwd_smooth_ms else {
                    "flow": "8x-downsampled N(0, 2^2) noise, bilinearly upsampled (SURVEY.md 8d 'typical'; what model.py:78 produces)",
                    "avg_launch_ms_alone": fwd_smooth_ms,
                    "frac": fwd_bytes(*SHAPES["level2"]) / fwd_smooth_ms / 1e6 / peak}),
                "algorithmic_bytes_per_launch": fwd_bytes(B, C, H, W), "avg_launch_ms": fwd_l2_ms,
                "timing": roofline_timing, "avg_launch_ms_alone": fwd_alone_ms,
                "frac_of_nominal_8000": ach / 8000.0,
            },
            "cpu_baseline": cpu,
            "extras": extras,
        }
        emit(line)
    if world > 1:
        dist.destroy_process_group()
    return 0


# ------------------------------------------------------------------------------------------------
# whole-network legs (ALL ranks): BASELINE.json configs 3, 4 and 5 around the hot path
# ------------------------------------------------------------------------------------------------
FULL_PYRAMID_LEGS = [   # name, pairs per GPU, H, W, replayed from a CUDA graph
    ("448x384_B1", 1, 384, 448, True),        # config 1 shape on the GPU: batch-1 latency
    ("448x384_B64", 64, 384, 448, False),     # the north star's "pairs/s @448x384"
    ("384x512_B64", 64, 384, 512, False),     # config 3: FlyingChairs-shaped, batch 64 per GPU
    ("1024x448_B1", 1, 448, 1024, True),      # config 5: Sintel-shaped latency / throughput
    ("1024x448_B16", 16, 448, 1024, False),
    ("1280x384_B1", 1, 384, 1280, True),      # config 5: KITTI 1242x375 padded to 1280x384
    ("1280x384_B16", 16, 384, 1280, False),
]


def measure_network_legs(torch, dist, dev, rank, world, dump_kernels=None):
    """Config 4 (DDP training step, the project's only collective) and configs 3/5 (full-pyramid
    inference) with the fused operator in the loop.  Every rank runs them on its own image pairs; a
    figure is all ranks' pairs divided by the slowest rank's device time.  Each leg carries a same-run
    parity assertion against the torch oracle operators (the checker, never the thing timed)."""
    from pwc_net_pytorch_b200 import parallel
    from pwc_net_pytorch_b200.workloads import PyramidInference, TrainStep, multiscale_l1
    from pwc_net_pytorch_b200.model import Net, default_args
    out = {}

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, iters, warm):
        for _ in range(warm):
            fn()
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(iters):
            fn()
        e1.record()
        barrier()
        return parallel.max_over_ranks(e0.elapsed_time(e1) / iters, dev)

    tf32 = bool(torch.backends.cudnn.allow_tf32)

    # ---- parity of the training path at a small shape: fused operator vs torch oracle operators ----
    parity = {}
    try:
        from oracle.model_ops import TorchCorrelationOps, deterministic_init
        torch.backends.cudnn.allow_tf32 = False
        a = default_args(device=dev)
        nets = (Net(a).train(), Net(a, ops=TorchCorrelationOps(4)).train())
        g = torch.Generator().manual_seed(5)
        x = (torch.rand(2, 3, 2, 128, 192, generator=g) * 255.0).to(dev)
        gt = (torch.randn(2, 2, 128, 192, generator=g) * 3.0).to(dev)
        losses, grads = [], []
        for net in nets:
            deterministic_init(net, seed=4)
            flows, _ = net(x)
            loss = multiscale_l1(flows, gt)
            net.zero_grad()
            loss.backward()
            losses.append(float(loss.detach()))
            grads.append({k: p.grad for k, p in net.named_parameters() if p.grad is not None})
        gerr = max(float((grads[0][k] - grads[1][k]).abs().max() / grads[1][k].abs().max().clamp_min(1e-30))
                   for k in grads[1])
        parity = {"loss_rel_diff": abs(losses[0] - losses[1]) / abs(losses[1]), "param_grad_max_rel_diff": gerr,
                  "shape": "2 pairs 128x192, fp32 convs, fused CUDA op vs torch oracle ops in the same network"}
        parity["ok"] = bool(parity["loss_rel_diff"] <= 1e-6 and gerr <= 1e-2)
        del nets, grads
    except Exception as e:
        parity = {"error": repr(e)}
    finally:
        torch.backends.cudnn.allow_tf32 = tf32

    # ---- config 4: training step, batch 8 per GPU, 384x448 ----
    train = {"workload": "PWC-Net training step: fwd + MultiScale L1 (losses.py:80-98) + bwd through the fused "
                         "warp/corr op + Adam(lr 1e-4, wd 4e-4); 8 pairs/GPU 384x448; DDP gradient all-reduce "
                         "over NCCL for N > 1",
             "tf32_convs": tf32, "hot_path_dtype": "f32", "parity_small_shape": parity}
    try:
        for unused, cl in (("find", False), ("freeze", False), ("freeze", True)):
            ts = TrainStep(dev, batch=8, height=384, width=448, unused=unused, channels_last=cl)
            first = float(ts.step())
            ms = timed(ts.step, iters=8, warm=5)      # (cuDNN picks / compiles kernels during the first steps: 2 warm-ups once gave 45 instead of 31 ms)
            row = {"ms_per_step": ms, "pairs_per_s": world * ts.batch / (ms * 1e-3),
                   "allreduce_bytes": ts.grad_bytes(), "bucket_cap_mb": ts.bucket_cap_mb,
                   "loss_first": first, "loss_last": float(ts.step())}
            if world > 1:
                flat = torch.empty(ts.grad_bytes() // 4, device=dev)
                ms_ar = timed(lambda: dist.all_reduce(flat), iters=10, warm=3)
                row.update({"allreduce_alone_ms": ms_ar,
                            "allreduce_alone_GBps_busbw": ts.grad_bytes() * 2 * (world - 1) / world / ms_ar / 1e6})
                if unused == "freeze":
                    # what the collective costs inside the step: the same step with DDP.no_sync() (no all-reduce).
                    # Not taken with find_unused_parameters=True: there no_sync changes the reducer's bookkeeping
                    # as well and the comparison is not like for like (measured: the no_sync step is SLOWER).
                    ms_local = timed(lambda: ts.step(sync=False), iters=6, warm=1)
                    row.update({"ms_per_step_without_allreduce": ms_local, "allreduce_exposed_ms": ms - ms_local})
                if unused == "find":
                    row["nccl_kernels"] = profile_nccl(torch, ts, rank, dump_kernels)
            key = "find_unused_parameters" if unused == "find" else ("lv5_lv6_frozen_channels_last" if cl else "lv5_lv6_frozen")
            train[key] = row
            del ts
            torch.cuda.empty_cache()
        # headline of the leg: the reference's modules as they are (unused estimators found by DDP, NCHW weights);
        # the two policy variants (frozen unused estimators; channels_last weights) are reported beside it
        train["pairs_per_s"] = train["find_unused_parameters"]["pairs_per_s"]
        train["pairs_per_s_best_policy"] = max(v["pairs_per_s"] for k, v in train.items()
                                               if isinstance(v, dict) and "pairs_per_s" in v)
    except Exception as e:
        train["error"] = repr(e)
    out["train_step_ddp"] = train

    # ---- configs 3 and 5 (+ the 448x384 figures): full-pyramid inference ----
    torch.manual_seed(0)
    net = Net(default_args(device=dev)).eval()
    # conv-stack policy (the convolutions are cuDNN, out of scope as kernels): TF32
    # math with the weights in NCHW (torch default) and in channels_last (cuDNN's NHWC kernels: measured 1.4x
    # on this network); the hot path is fp32 NCHW in both.  `pairs_per_s` is the better of the two.
    for name, B, H, W, graphed in FULL_PYRAMID_LEGS:
        row = {"pairs_per_gpu": B, "launch": "cuda_graph" if graphed else "eager", "tf32_convs": tf32}
        try:
            for fmt_name, fmt in (("nchw", torch.contiguous_format), ("channels_last", torch.channels_last)):
                net.to(memory_format=fmt)
                leg = PyramidInference(dev, B, H, W, graphed=graphed, net=net)
                ms = timed(leg.run, iters=5 if B > 1 else 20, warm=2)
                row[f"ms_{fmt_name}"] = ms
                row[f"pairs_per_s_{fmt_name}"] = world * B / (ms * 1e-3)
                del leg
                torch.cuda.empty_cache()
            row["ms"] = min(row["ms_nchw"], row["ms_channels_last"])
            row["pairs_per_s"] = world * B / (row["ms"] * 1e-3)
        except Exception as e:
            row["error"] = repr(e)
        out[f"full_pyramid_{name}"] = row
    net.to(memory_format=torch.contiguous_format)
    # same-run parity: end-point-error delta of the full forward, fused CUDA op vs torch oracle ops, fp32 convs
    try:
        from oracle.model_ops import TorchCorrelationOps
        torch.backends.cudnn.allow_tf32 = False
        oracle_net = Net(default_args(device=dev), ops=TorchCorrelationOps(4)).eval()
        oracle_net.load_state_dict(net.state_dict())
        for name, H, W in (("384x512", 384, 512), ("1024x448", 448, 1024), ("1280x384", 384, 1280)):
            g = torch.Generator().manual_seed(H + W)
            x = (torch.rand(1, 3, 2, H, W, generator=g) * 255.0).to(dev)
            with torch.no_grad():
                fa, _ = net(x)
                fb, _ = oracle_net(x)
            epe = max(float(torch.norm(a - b, p=2, dim=1).max()) for a, b in zip(fa, fb))
            out[f"full_pyramid_epe_delta_px_{name}"] = {"max_epe_delta_px": epe, "ok": bool(epe <= 1e-4),
                                                        "bound": 1e-4}
        del oracle_net
    except Exception as e:
        out["full_pyramid_epe_error"] = repr(e)
    finally:
        torch.backends.cudnn.allow_tf32 = tf32
    del net
    torch.cuda.empty_cache()
    return out


def profile_nccl(torch, ts, rank, dump_path=None):
    """Device time of the NCCL kernels inside the DDP step, from the CUPTI kernel trace of three steps
    (torch.profiler).  Returns None when the trace is unavailable."""
    try:
        from torch.profiler import ProfilerActivity, profile
        ts.step()
        torch.cuda.synchronize()
        with profile(activities=[ProfilerActivity.CUDA]) as prof:
            for _ in range(3):
                ts.step()
            torch.cuda.synchronize()
        rows = {}
        for ev in prof.events():
            if getattr(ev, "device_type", None) is not None and "cuda" in str(ev.device_type).lower():
                r = rows.setdefault(ev.name, [0, 0.0])
                r[0] += 1
                r[1] += float(getattr(ev, "device_time", 0.0) or getattr(ev, "cuda_time", 0.0))
        total = sum(v[1] for v in rows.values())
        nccl = sum(v[1] for k, v in rows.items() if "nccl" in k.lower())
        if dump_path and rank == 0:
            with open(dump_path, "w") as f:
                f.write("# CUDA kernels of 3 DDP training steps (torch.profiler / CUPTI), rank 0: name, launches, total us\n")
                for k, v in sorted(rows.items(), key=lambda kv: -kv[1][1]):
                    f.write(f"{v[1]:12.1f} us {v[0]:6d} x  {k[:160]}\n")
        return {"ms_per_step_incl_peer_wait": nccl / 3e3, "share_of_kernel_time": (nccl / total) if total else None,
                "kernel_ms_per_step": total / 3e3,
                "note": "duration of the NCCL all-reduce kernel from CUPTI: it is launched when this rank's bucket is "
                        "ready and spins until the peer arrives, so it measures rank skew, not transfer time "
                        "(transfer alone: allreduce_alone_ms)"}
    except Exception as e:
        return {"error": repr(e)}


def time_cuda(torch, fn, iters=20, warm=3):
    """Mean device time of fn() over `iters` back-to-back calls.  A device-side sleep is queued first so
    that the launches are already enqueued when the GPU reaches them: short kernels (coarse pyramid
    levels, 10-20 us) are then not measured at the pace of the Python launch path."""
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda._sleep(int(2.0e6 * max(1, iters // 10)))      # ~1 ms per 10 queued calls
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


SM_COUNT_B200 = 148


def measure_extras(torch, pkg, dev, sets, make_set):
    """Per-kernel timings (CUDA events, rotating inputs) that explain the headline."""
    from pwc_net_pytorch_b200 import functional as PF
    peaks, _ = measured_peaks()
    peak = float(peaks.get("hbm_gbs", 6650.0))
    out = {"kernels": {}}
    op = pkg.FusedWarpCorrelation(**CANON)
    op_ref = pkg.FusedWarpCorrelation.from_search_range(4)
    smooth = [make_set("smooth") for _ in range(3)]
    ctr = [0]

    def rot(lst):
        ctr[0] += 1
        return lst[ctr[0] % len(lst)]

    for name, (B, C, H, W) in SHAPES.items():
        fb, bb = fwd_bytes(B, C, H, W), bwd_bytes(B, C, H, W)
        with torch.no_grad():
            ms = time_cuda(torch, lambda: op(*rot(sets)[name][:3]))
            out["kernels"][f"fwd_{name}_iid"] = {"ms": ms, "GBps": fb / ms / 1e6, "frac": fb / ms / 1e6 / peak}
            ms = time_cuda(torch, lambda: op(*rot(smooth)[name][:3]))
            out["kernels"][f"fwd_{name}_smooth"] = {"ms": ms, "GBps": fb / ms / 1e6, "frac": fb / ms / 1e6 / peak}
            ms = time_cuda(torch, lambda: op_ref(*rot(sets)[name][:3]))
            out["kernels"][f"fwd_{name}_iid_refcfg_s2"] = {"ms": ms, "GBps": fb / ms / 1e6, "frac": fb / ms / 1e6 / peak}

        def fb_step():
            f1, f2, flow, gout = rot(sets)[name]
            a, b, f = f1.requires_grad_(), f2.requires_grad_(), flow.requires_grad_()
            a.grad = b.grad = f.grad = None
            op(a, b, f).backward(gout)
        ms_fb = time_cuda(torch, fb_step)
        ms_f = out["kernels"][f"fwd_{name}_iid"]["ms"]
        ms_b = max(ms_fb - ms_f, 1e-6)
        out["kernels"][f"bwd_{name}_iid"] = {"ms": ms_b, "GBps": bb / ms_b / 1e6, "frac": bb / ms_b / 1e6 / peak,
                                            "note": "fwd+bwd minus fwd; includes autograd overhead"}

    # the five fused calls of 384x448 pairs, forward, batch 32
    B = 32
    lv = []
    for (C, H, W) in PYRAMID_384x448:
        lv.append((torch.randn(B, C, H, W, device=dev), torch.randn(B, C, H, W, device=dev),
                   2.0 * torch.randn(B, 2, H, W, device=dev)))
    with torch.no_grad():
        def pyr():
            for a, b, f in lv:
                op(a, b, f)
        ms = time_cuda(torch, pyr)
    pb = sum(fwd_bytes(B, C, H, W) for (C, H, W) in PYRAMID_384x448)
    out["pyramid5_fwd_B32_384x448"] = {"ms": ms, "pairs_per_s": B / (ms * 1e-3), "GBps": pb / ms / 1e6,
                                       "frac": pb / ms / 1e6 / peak}
    # the honest per-pair figure of the hot path: all five levels, forward + backward, batch 32
    for a, b, f in lv:
        a.requires_grad_(); b.requires_grad_(); f.requires_grad_()
    gos = [torch.randn(B, 81, H, W, device=dev) for (C, H, W) in PYRAMID_384x448]

    def pyr_fb():
        for (a, b, f), go in zip(lv, gos):
            a.grad = b.grad = f.grad = None
            op(a, b, f).backward(go)
    ms = time_cuda(torch, pyr_fb, iters=10, warm=2)
    pbb = pb + sum(bwd_bytes(B, C, H, W) for (C, H, W) in PYRAMID_384x448)
    out["pyramid5_fwdbwd_B32"] = {"ms": ms, "pairs_per_s": B / (ms * 1e-3), "GBps": pbb / ms / 1e6,
                                  "frac": pbb / ms / 1e6 / peak,
                                  "note": "five fused calls of 32 pairs at 384x448, fwd+bwd incl. autograd overhead"}
    del lv, gos
    # per-level table (north_star: fraction of the HBM roofline at pyramid levels 2-6), forward and
    # backward, literal reference configuration (pad 9 / md 9 / stride2 2) and canonical md=4
    levels = {}
    shapes = [(f"L{6 - i}_448x384", B, C, H, W) for i, (C, H, W) in enumerate(PYRAMID_384x448)]
    shapes += [("L2_sintel_1024x448", 16, 32, 112, 256), ("L2_kitti_1280x384", 16, 32, 96, 320),
               ("L2_chairs_512x384_B64", 64, 32, 96, 128)]
    for name, Bl, C, H, W in shapes:
        a = torch.randn(Bl, C, H, W, device=dev)
        b = torch.randn(Bl, C, H, W, device=dev)
        f = 2.0 * torch.randn(Bl, 2, H, W, device=dev)
        go = torch.randn(Bl, 81, H, W, device=dev)
        row = {"B": Bl, "C": C, "H": H, "W": W}
        for tag, o in (("canon", op), ("refcfg", op_ref)):
            with torch.no_grad():
                ms_f = time_cuda(torch, lambda: o(a, b, f), iters=10, warm=2)
            a.requires_grad_(); b.requires_grad_(); f.requires_grad_()

            def fb():
                a.grad = b.grad = f.grad = None
                o(a, b, f).backward(go)
            ms_fb = time_cuda(torch, fb, iters=10, warm=2)
            a.requires_grad_(False); b.requires_grad_(False); f.requires_grad_(False)
            fbytes, bbytes = fwd_bytes(Bl, C, H, W), bwd_bytes(Bl, C, H, W)
            row[tag] = {"fwd_ms": ms_f, "fwd_frac": fbytes / ms_f / 1e6 / peak,
                        "bwd_ms": max(ms_fb - ms_f, 1e-6), "bwd_frac": bbytes / max(ms_fb - ms_f, 1e-6) / 1e6 / peak}
        # the two floors of the forward: HBM at the measured copy bandwidth, and the fp32 FMA pipe -- at
        # its 128 lane-FMA/clk/SM peak and at the ~57 % an 81-displacement band correlation sustains in a SIMT register
        # tiling (profiles/r02_forward_limits.txt: FFMA 1.48 clk in an outer-product pattern, LDS.128 4 clk)
        fma = 81.0 * C * Bl * H * W
        hbm_ms = fwd_bytes(Bl, C, H, W) / (peak * 1e6)
        fma_ms_peak = fma / (SM_COUNT_B200 * 128 * 1.965e9) * 1e3
        row["fwd_floors_ms"] = {"hbm": hbm_ms, "fma_peak": fma_ms_peak, "fma_sustained_simt": fma_ms_peak / 0.57,
                                "bound": "fma" if fma_ms_peak / 0.57 > hbm_ms else "hbm"}
        levels[name] = row
    out["levels"] = levels

    # full PWC-Net forward (reference architecture, random init, convs on cuDNN, hot path fused):
    # the north star's "pairs/s at 448x384" figure; the conv stack is out of scope
    try:
        from pwc_net_pytorch_b200.model import Net, default_args
        net = Net(default_args(device=dev)).eval()
        for tf32 in (True, False):
            torch.backends.cudnn.allow_tf32 = tf32
            for Bf in (1, 16):
                xin = torch.rand(Bf, 3, 2, 384, 448, device=dev) * 255.0
                with torch.no_grad():
                    ms = time_cuda(torch, lambda: net(xin), iters=5, warm=2)
                out[f"full_forward_384x448_B{Bf}_{'tf32' if tf32 else 'fp32'}convs"] = {
                    "ms": ms, "pairs_per_s": Bf / (ms * 1e-3)}
        torch.backends.cudnn.allow_tf32 = True
        # reduced-precision convolutions (bf16 autocast, channels-last): NOT within the 1e-4 px parity bound
        # (random-init weights: ~0.03 px EPE from the TF32 result) -- an indication of what the conv stack,
        # which is out of scope here, leaves on the table; the hot path itself stays fp32
        net_cl = net.to(memory_format=torch.channels_last)
        for Bf in (16, 64):
            xin = torch.rand(Bf, 3, 2, 384, 448, device=dev) * 255.0
            with torch.no_grad(), torch.autocast("cuda", dtype=torch.bfloat16):
                ms = time_cuda(torch, lambda: net_cl(xin), iters=5, warm=2)
            out[f"full_forward_384x448_B{Bf}_bf16convs_channels_last"] = {"ms": ms, "pairs_per_s": Bf / (ms * 1e-3)}
        net = net.to(memory_format=torch.contiguous_format)
        # whole forward replayed from one CUDA graph (row f: the coarse levels are launch latency)
        from pwc_net_pytorch_b200.graphed import GraphedForward
        for Bf in (1, 16):
            xin = torch.rand(Bf, 3, 2, 384, 448, device=dev) * 255.0
            gf = GraphedForward(net, xin)
            ms = time_cuda(torch, lambda: gf(xin), iters=10, warm=2)
            out[f"full_forward_384x448_B{Bf}_tf32convs_cudagraph"] = {"ms": ms, "pairs_per_s": Bf / (ms * 1e-3)}
            del gf
        del net
    except Exception as e:
        out["full_forward_error"] = repr(e)

    # GPU reference bar: the reference's own kernels (sm_100a build) + torch grid_sample
    try:
        from oracle import ref_cuda
        from oracle import torch_ref as tr
        if ref_cuda.available():
            for name, (B, C, H, W) in SHAPES.items():
                f1, f2, flow, gout = sets[0][name]
                with torch.no_grad():
                    ms = time_cuda(torch, lambda: ref_cuda.correlation_forward(
                        f1, tr.warping_layer_port(f2, flow), 4, 1, 4, 1, 1), iters=5, warm=2)
                out["kernels"][f"gpu_reference_fwd_{name}"] = {
                    "ms": ms, "note": "reference correlation_cuda_kernel.cu (unchanged, sm_100a) + its fills + "
                                      "torch grid_sample WarpingLayer port"}

                # forward + backward the way the reference runs it on a GPU: its own kernels for the
                # correlation and its gradients, torch autograd (grid_sample backward) for the warp
                def ref_fb():
                    x2 = f2.detach().requires_grad_()
                    fl = flow.detach().requires_grad_()
                    w = tr.warping_layer_port(x2, fl)
                    with torch.no_grad():
                        ref_cuda.correlation_forward(f1, w, 4, 1, 4, 1, 1)
                        _, g2 = ref_cuda.correlation_backward(gout, f1, w.detach(), 4, 1, 4, 1, 1)
                    w.backward(g2)
                ms_fb = time_cuda(torch, ref_fb, iters=3, warm=1)
                out["kernels"][f"gpu_reference_fwdbwd_{name}"] = {
                    "ms": ms_fb, "pairs_per_s": B / (ms_fb * 1e-3),
                    "note": "the reference's default GPU path on this B200: its .cu unchanged (2*B backward "
                            "launches) + torch grid_sample fwd/bwd"}
            k = out["kernels"]
            ref_ms = sum(k[f"gpu_reference_fwdbwd_{n}"]["ms"] for n in SHAPES)
            own_ms = sum(k[f"fwd_{n}_iid"]["ms"] + k[f"bwd_{n}_iid"]["ms"] for n in SHAPES)
            out["vs_gpu_reference"] = {
                "reference_ms_per_step": ref_ms, "native_ms_per_step_serial": own_ms, "speedup": ref_ms / own_ms,
                "note": "cfg2 step (both shapes, fwd+bwd, 32 pairs) on the same B200: reference CUDA kernels + torch "
                        "grid_sample vs this library, both device-resident and timed back to back on one stream"}
    except Exception as e:   # the bar is optional evidence, never the product path
        out["gpu_reference_error"] = repr(e)
    return out


_REAL_STDOUT = None


def emit(line):
    """Writes the one JSON line to the process's original stdout."""
    data = (json.dumps(line) + "\n").encode()
    if _REAL_STDOUT is None:
        sys.stdout.write(data.decode())
        sys.stdout.flush()
    else:
        os.write(_REAL_STDOUT, data)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", choices=["native", "reference"], default="native")
    ap.add_argument("--no-extras", action="store_true")
    ap.add_argument("--eager", action="store_true",
                    help="launch every kernel from Python instead of replaying the step from CUDA graphs (the "
                         "default; every entry point of the library is capturable: no allocation, no sync)")
    ap.add_argument("--graph", action="store_true", help="(default; kept for compatibility)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--dump-kernels", default=None,
                    help="write the CUPTI kernel list of the DDP training step (rank 0, N > 1) to this file")
    args = ap.parse_args()
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if not (args.impl == "native" and args.gpus > 1 and world == 1):
        # stdout carries exactly ONE JSON line: every other writer to file descriptor 1 (NCCL's version banner,
        # library chatter) is sent to stderr at the descriptor level; the line goes to the saved descriptor
        global _REAL_STDOUT
        sys.stdout.flush()
        _REAL_STDOUT = os.dup(1)
        os.dup2(2, 1)
    if args.impl == "reference":
        return run_reference(args)
    if args.gpus > 1 and world == 1:
        # convenience: re-launch under torchrun
        import subprocess
        cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={args.gpus}",
               "--master-addr", "127.0.0.1", "--master-port", os.environ.get("MASTER_PORT", "29511"),
               os.path.abspath(__file__)] + sys.argv[1:]
        return subprocess.call(cmd)
    return run_native(args)


if __name__ == "__main__":
    sys.exit(main())
