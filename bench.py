#!/usr/bin/env python
"""Benchmark of the Fourier-layer hot path: FNO train samples/sec (fwd + bwd + Adam).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload cfg5] [--impl reference]

One "step" = zero_grad -> forward -> rel-L2 sum loss -> backward -> (SUM all-reduce, N > 1) -> Adam on
one synthetic batch (reference loop: advection_diffusion/func_to_func/train_fno.py:187-197).
Default workload = BASELINE.json configs[4] ("large FNO2d 256x256, modes 32, width 128"), the
config the 1/2/4/8-GPU metric is quoted on, at its per-GPU batch of 32 (global 256 on 8 GPUs, weak
scaling).  Other BASELINE configs: --workload cfg1 | cfg2 | cfg3 | cfg4.

Prints ONE JSON line (rank 0).  `value` is device-timed with inputs resident in HBM; `e2e` includes the
pinned-host -> device copy of each step's batch and the device -> host read of the loss.
`--impl reference` times the reference algorithm's CPU path (the oracle port: same torch.fft /
einsum / conv calls as the reference, on all host cores) on a bounded sample of the same workload.
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

WORKLOADS = {
    # name: (description, family, ctor kwargs, per-GPU batch, x shape (no batch), CPU sample batch)
    "cfg1": ("FNO2d 64x64 modes12 width32 L4 batch20", "fno2d",
             dict(modes1=12, modes2=12, width=32), 20, (1, 64, 64), 20),
    "cfg2": ("FNO1d 1024 modes16 width64 L4 batch64", "fno1d",
             dict(modes1=16, width=64), 64, (1, 1024), 64),
    "cfg3": ("FNF2d 221x51 modes12 width32 batch128", "fnf2d",
             dict(modes1=12, modes2=12, width=32, d_in=2, d_out=2), 128, (2, 221, 51), 16),
    "cfg4": ("FND2d 20->128x128 modes16 width64 batch20", "fnd2d",
             dict(d_in=20, s_outputspace=(128, 128), modes1=16, modes2=16, width=64), 20, (20,), 4),
    "cfg5": ("FNO2d 256x256 modes32 width128 L4 batch32/GPU (global 256 on 8 GPUs)", "fno2d",
             dict(modes1=32, modes2=32, width=128), 32, (1, 256, 256), 1),
}


def build_model(family, kw):
    import fnm_b200
    ctor = {"fno2d": fnm_b200.FNO2d, "fno1d": fnm_b200.FNO1d, "fnf2d": fnm_b200.FNF2d, "fnd2d": fnm_b200.FND2d}[family]
    return ctor(**kw)


def oracle_forward(family, kw):
    from oracle import fnm_oracle as O
    if family in ("fno2d", "fno1d"):
        return lambda p, x: O.f2f_forward(p, x)
    if family == "fnf2d":
        return lambda p, x: O.f2v_forward(p, x)
    return lambda p, x: O.v2f_forward(p, x, kw["s_outputspace"])


def synth(shape, seed):
    return torch.randn(*shape, generator=torch.Generator().manual_seed(seed))


# ------------------------------------------------------------------------------------------------
class ClockSampler:
    """SM clocks / throttle reasons sampled DURING the timed region: NVML in-process every ~2 ms (the small workloads'
    timed region is ~10 ms, far below what an `nvidia-smi -lms` subprocess resolves), nvidia-smi as the fallback."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")
    BITS = {0x4: "sw_power_cap", 0x8: "hw_slowdown", 0x20: "sw_thermal_slowdown", 0x40: "hw_thermal_slowdown"}

    def __init__(self, index):
        self.index, self.proc, self.lines = index, None, []
        self.nvml, self.handle, self.samples, self.mask, self.mx, self.run = None, None, [], 0, None, False
        try:
            import pynvml
            pynvml.nvmlInit()
            try:
                uuid = str(torch.cuda.get_device_properties(index).uuid)
                self.handle = pynvml.nvmlDeviceGetHandleByUUID(("GPU-" + uuid) if not uuid.startswith("GPU-") else uuid)
            except Exception:                              # noqa: BLE001
                self.handle = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.mx = float(pynvml.nvmlDeviceGetMaxClockInfo(self.handle, pynvml.NVML_CLOCK_SM))
            self.nvml = pynvml
        except Exception:                                  # noqa: BLE001 -- no NVML: nvidia-smi below
            self.nvml = None

    def _poll(self):
        nv = self.nvml
        reasons = getattr(nv, "nvmlDeviceGetCurrentClocksEventReasons", None) or nv.nvmlDeviceGetCurrentClocksThrottleReasons
        while self.run:
            try:
                self.samples.append(float(nv.nvmlDeviceGetClockInfo(self.handle, nv.NVML_CLOCK_SM)))
                self.mask |= int(reasons(self.handle))
            except Exception:                              # noqa: BLE001
                pass
            time.sleep(0.002)

    def start(self):
        if self.nvml is not None:
            self.run = True
            self.thread = threading.Thread(target=self._poll, daemon=True)
            self.thread.start()
            return
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100"],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.nvml is not None:
            self.run = False
            self.thread.join(timeout=1)
            sm = self.samples
            return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": self.mx, "samples": len(sm),
                    "reasons": sorted(n for b, n in self.BITS.items() if self.mask & b), "source": "nvml"}
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            f = [c.strip() for c in ln.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0]))
                mx.append(float(f[1]))
            except ValueError:
                continue
            for n, v in zip(names, f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "samples": len(sm), "reasons": sorted(reasons), "source": "nvidia-smi"}


def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        d = json.load(open(path))
        return d.get("hbm_gbs", 6650.0), "MEASURED_PEAKS.json hbm_gbs"
    return 6650.0, "fallback 6.65 TB/s (B200_PROFILING.md)"


# ------------------------------------------------------------------------------------------------
def _oracle_setup(workload, device="cpu"):
    from oracle import fnm_oracle as O
    desc, family, kw, batch, xshape, cpu_batch = WORKLOADS[workload]
    torch.manual_seed(0)
    model = build_model(family, kw)
    p = {k: v.detach().clone().to(device).requires_grad_(True) for k, v in model.state_dict().items()}
    del model
    return O, desc, family, kw, batch, xshape, cpu_batch, p, oracle_forward(family, kw)


def _cpu_pick_batch(O, fwd, p, xshape, batch, start, budget_s, steps_wanted):
    """Largest batch <= the workload's whose `steps_wanted` steps fit `budget_s`, assuming time linear in the batch
    (measured with one probe step at `start`; the probe is also the allocator / thread-pool warm-up)."""
    b0 = max(1, min(start, batch))
    x = synth((b0,) + xshape, 1234)
    with torch.no_grad():
        y = synth(tuple(fwd(p, x).shape), 1235)
    O.train_step(fwd, p, x, y, O.AdamState(), lr=1e-3, weight_decay=1e-4)          # first call: page-in, not timed
    t0 = time.perf_counter()
    O.train_step(fwd, p, x, y, O.AdamState(), lr=1e-3, weight_decay=1e-4)
    per_sample = (time.perf_counter() - t0) / b0
    b = int(budget_s / max(per_sample * steps_wanted, 1e-9))
    return max(1, min(batch, b))


def run_reference(args, rank):
    """CPU arm: the oracle port of the reference algorithm (the same torch.fft / einsum / conv CPU entry points the
    reference calls) on all host cores, on a bounded sample of the same workload: the largest batch (<= the workload's
    own) whose K + W steps fit ~100 s."""
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    O, desc, family, kw, batch, xshape, cpu_batch, p, fwd = _oracle_setup(args.workload)
    cpu_batch = _cpu_pick_batch(O, fwd, p, xshape, batch, 2 if batch >= 2 else 1, 100.0, args.steps + args.warmup)
    x = synth((cpu_batch,) + xshape, 1234)
    with torch.no_grad():
        yshape = fwd(p, x).shape
    y = synth(tuple(yshape), 1235)
    state = O.AdamState()
    for _ in range(args.warmup):
        O.train_step(fwd, p, x, y, state, lr=1e-3, weight_decay=1e-4)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        O.train_step(fwd, p, x, y, state, lr=1e-3, weight_decay=1e-4)
    dt = time.perf_counter() - t0
    value = cpu_batch * args.steps / dt
    sample = (f"batch {cpu_batch} of the workload's {batch} per step (largest batch whose {args.steps}+{args.warmup} steps fit "
              f"100 s of CPU time), {args.steps} steps after {args.warmup} warm-up")
    print(json.dumps({
        "impl": "reference", "metric": "train_samples_per_sec", "value": value, "unit": "samples/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": f"{args.workload}: {desc}", "device": "cpu", "threads": cores, "cpu_batch": cpu_batch},
        "cpu_baseline": {"value": value, "unit": "samples/s", "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": "samples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }))


def cpu_baseline(workload, budget_s=25.0):
    """Bounded CPU sample of the same step with the oracle port (reported baseline, not a target): the largest batch
    whose 1 warm-up + 2 timed steps fit the budget."""
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    O, desc, family, kw, batch, xshape, cpu_batch, p, fwd = _oracle_setup(workload)
    cpu_batch = _cpu_pick_batch(O, fwd, p, xshape, batch, 2 if batch >= 2 else 1, budget_s, 3)
    x = synth((cpu_batch,) + xshape, 1234)
    with torch.no_grad():
        y = synth(tuple(fwd(p, x).shape), 1235)
    state = O.AdamState()
    O.train_step(fwd, p, x, y, state, lr=1e-3, weight_decay=1e-4)           # warm-up
    n, t0 = 0, time.perf_counter()
    while True:
        O.train_step(fwd, p, x, y, state, lr=1e-3, weight_decay=1e-4)
        n += 1
        dt = time.perf_counter() - t0
        if dt > budget_s or n >= 20 or (n >= 2 and dt * (n + 1) / n > budget_s):
            break
    return {"value": cpu_batch * n / dt, "unit": "samples/s", "cores": cores, "kind": "port",
            "sample": f"batch {cpu_batch} of {batch} (largest that fits the {budget_s:.0f} s budget), {n} steps after 1 warm-up, {dt:.1f} s"}


# SMs reserved for (and CTAs granted to) the NCCL kernels of the overlapped gradient exchange, by world size: ring
# channels at 2 GPUs need more CTAs than the NVLS (in-switch) reduction at 4 and 8
DP_SMS = {2: 16, 4: 16, 8: 16}


def gpu_comparator(workload, steps, warmup, allow_tf32, dev):
    """The reference's OWN GPU path on this box: the oracle port -- the same torch.fft.rfft2 / irfft2
    (cuFFT), einsum (cuBLAS), 1x1 Conv (cuDNN), F.gelu and reference-semantics Adam calls as /root/reference/models/shared.py:
    315-341 and func_to_func2d.py:62-65,89-95 -- run on the GPU through stock PyTorch, same batch, loss and optimizer."""
    O, desc, family, kw, batch, xshape, _, p, fwd = _oracle_setup(workload, dev)
    old = (torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32)
    torch.backends.cudnn.allow_tf32 = torch.backends.cuda.matmul.allow_tf32 = bool(allow_tf32)
    try:
        x = synth((batch,) + xshape, 1234).to(dev)
        with torch.no_grad():
            y = synth(tuple(fwd(p, x).shape), 1235).to(dev)
        state = O.AdamState()
        for _ in range(max(warmup, 3)):
            O.train_step(fwd, p, x, y, state, lr=1e-3, weight_decay=1e-4)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            loss, _ = O.train_step(fwd, p, x, y, state, lr=1e-3, weight_decay=1e-4)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / steps
        return {"value": batch / (ms / 1e3), "unit": "samples/s", "ms_per_step": ms, "batch": batch, "steps": steps,
                "allow_tf32": bool(allow_tf32), "last_loss": float(loss),
                "what": "stock PyTorch on the same GPU: cuFFT rfft2/irfft2 + cuBLAS einsum + cuDNN 1x1 conv + reference Adam"}
    except torch.cuda.OutOfMemoryError as e:                       # noqa: PERF203
        return {"unavailable": f"out of memory: {str(e)[:120]}"}
    finally:
        torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32 = old
        del p
        torch.cuda.empty_cache()


def run_stock_gpu(args, rank):
    """`--impl stock-gpu`: only the comparator, as its own JSON line (rank 0)."""
    if rank != 0:
        return
    dev = torch.device("cuda", int(os.environ.get("LOCAL_RANK", 0)))
    torch.cuda.set_device(dev)
    desc, family, kw, batch, xshape, _ = WORKLOADS[args.workload]
    res = {tag: gpu_comparator(args.workload, args.steps, args.warmup, tf32, dev) for tag, tf32 in (("fp32", False), ("tf32", True))}
    main_ = res["fp32"]
    print(json.dumps({
        "impl": "stock-gpu", "metric": "train_samples_per_sec", "value": main_.get("value"), "unit": "samples/s", "n_gpus": 1,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": main_.get("ms_per_step"), "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": f"{args.workload}: {desc}", "batch_per_gpu": batch, "device": "cuda (stock PyTorch)"},
        "gpu_comparator": res,
    }))


# ------------------------------------------------------------------------------------------------
def kernel_algorithmic_bytes(name, wl, model, launches=None):
    """ALGORITHMIC bytes one train step moves through kernel `name`:
    A = 4*B*C*P1*P2 bytes is one fp32 activation of the trunk."""
    A, L, lyf = wl["A"], wl["layers"], wl["ly_frac"]
    return {
        # fwd: read x, write z (2A; the activated copy some launches also write is a design choice, not counted);
        # bwd: read g_z, z_prev, write g_in (3A; the first layer has no GELU' operand: 2A)
        "pointwise_ysyn": L * 2 * A + (L - 1) * 3 * A + 2 * A,
        "ydft": (launches or 2 * L) * A * (1 + lyf),    # fwd (+ bwd) analysis: read A, write the LY/P2 partial transform
        # backward analysis fused with the conv weight gradient: ONE read of g, one of x, the partial transform written
        "ydft_conv_wgrad": (launches or L) * A * (2 + lyf),
        "conv_wgrad": (launches or L) * 2 * A,          # read g and x (one more launch for the projection's fc1)
        "pointwise_conv": 2 * 2 * A,                    # projection fc1 forward + its input gradient
        "lift_fwd": A, "lift_bwd": A, "mlp_head_fwd": A, "mlp_head_bwd": 2 * A,
        "adam": 28 * wl["n_real_params"],               # read p,g,m,v; write p,m,v
        "xdft": 2 * L * A * lyf, "xsyn": 2 * L * A * lyf,
        "mix_fwd": L * wl["w_bytes"], "mix_bwd_x": L * wl["w_bytes"], "mix_bwd_w": L * wl["w_bytes"],
        "mix_pack": 2 * L * 2 * wl["w_bytes"], "mix_unpack": L * 2 * wl["w_bytes"],
    }.get(name)


def run_ours(args, rank, world):
    # NCCL's log level is the caller's to choose (the driver reads the communicator lines); only a default is set here,
    # and it is NOT silenced.  The JSON line is the last line of stdout either way.
    os.environ.setdefault("NCCL_DEBUG", "WARN")
    if world > 1:
        # The gradient exchange runs UNDER the backward pass: the persistent kernels leave DP_SMS[world] SMs free while an
        # all-reduce is in flight and NCCL is held to as many CTAs (measured with tools/dp_timeline.py, profiles/README.md).
        # Both are defaults the caller's environment overrides.
        k = os.environ.setdefault("FNM_DP_SM_RESERVE", str(DP_SMS.get(world, 12)))
        if int(k) > 0:
            os.environ.setdefault("NCCL_MAX_CTAS", k)
            os.environ.setdefault("NCCL_MAX_NCHANNELS", k)
            os.environ.setdefault("NCCL_NVLS_NCHANNELS", k)
    import torch.distributed as dist
    import fnm_b200
    from fnm_b200 import _lib
    from fnm_b200.parallel import FlatGradients, rel_l2_sum

    local = int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    desc, family, kw, batch, xshape, _ = WORKLOADS[args.workload]
    fnm_b200.set_compute_mode(args.mode)
    torch.manual_seed(0)
    model = build_model(family, kw).to(dev)
    fnm_b200.parallel.broadcast_parameters(model)
    opt = fnm_b200.Adam(model.parameters(), lr=1e-3, weight_decay=1e-4, capturable=args.graph)
    grads = FlatGradients(model.parameters(), direct_spectral=True,       # dW kernels write into the flat buffer
                          overlap=os.environ.get("FNM_BENCH_OVERLAP", "1") == "1")   # slots exchanged as backward fills them (N > 1)

    x_host = synth((batch,) + xshape, 1234 + rank).pin_memory()
    x = x_host.to(dev)
    with torch.no_grad():
        yshape = tuple(model(x).shape)
    y_host = synth(yshape, 1235 + rank).pin_memory()
    y = y_host.to(dev)

    # SMs the persistent backward kernels leave to the NCCL kernels of the overlapped exchange (N > 1 only)
    sm_reserve = int(os.environ.get("FNM_DP_SM_RESERVE", "0")) if world > 1 else 0

    def step(xd, yd):
        return fnm_b200.parallel.train_step(model, opt, grads, rel_l2_sum, xd, yd, sm_reserve=sm_reserve)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(ms):
        if world == 1:
            return ms
        t = torch.tensor([ms], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return t.item()

    eager_step = step
    graphed = None
    if args.graph:
        # the step recorded once in a CUDA graph; every timed step below is one replay of the full step
        try:
            graphed = fnm_b200.parallel.GraphedTrainStep(model, opt, grads, rel_l2_sum, x, y, warmup=max(3, args.warmup),   # capture warm-up (undone)
                                                             sm_reserve=sm_reserve)
        except Exception as e:                             # noqa: BLE001 -- e.g. a collective that cannot be captured
            sys.stderr.write(f"[bench] CUDA-graph capture failed ({type(e).__name__}: {e}); running eagerly\n")
            graphed, args.graph = None, False
    if graphed is not None:
        def step(xd, yd):                                  # noqa: F811
            return graphed(xd, yd)
    for _ in range(args.warmup):
        step(x, y)
    # ---- timed region: K steps, inputs resident in HBM
    barrier()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    launches0 = _lib.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        step(x, y)
    e1.record()
    barrier()
    ms = max_over_ranks(e0.elapsed_time(e1))
    launches = _lib.launch_count() - launches0
    if graphed is not None:
        launches = graphed.launches_per_step * args.steps    # kernels recorded in the graph, replayed K times
    clocks = sampler.stop() if rank == 0 else None

    # ---- end to end: host batch -> device, step, loss -> host, every step
    xd, yd = torch.empty_like(x), torch.empty_like(y)
    # Input pipeline of the graphed step: every step's batch is copied from pinned host memory inside the timed region, the
    # copy of step i + 1 on a side stream while step i runs (GraphedTrainStep.stage); the loss is read back every step.
    loss_host = [torch.empty(1, dtype=torch.float32).pin_memory() for _ in range(2)]
    loss_ev = [torch.cuda.Event(), torch.cuda.Event()]

    def e2e_loop(n):
        last = None
        if graphed is not None and os.environ.get("FNM_BENCH_E2E_SIMPLE") == "1":   # copy, replay, loss.item() in sequence
            for _ in range(n):
                last = graphed(x_host, y_host).item()
            return last
        if graphed is not None:
            graphed.stage(x_host, y_host)                  # step 0's inputs: nothing to hide this copy behind
            for i in range(n):
                loss = graphed()
                loss_host[i % 2].copy_(loss.detach().reshape(1), non_blocking=True)      # device -> pinned host, every step
                loss_ev[i % 2].record()
                if i + 1 < n:
                    graphed.stage(x_host, y_host)          # the next step's 16.8 MB travel under this step's kernels
                if i > 0:                                  # the host runs ONE step ahead: it reads the loss of step i - 1
                    loss_ev[(i - 1) % 2].synchronize()     # while step i is on the GPU (no idle gap at the read-back)
                    last = float(loss_host[(i - 1) % 2])
            loss_ev[(n - 1) % 2].synchronize()
            return float(loss_host[(n - 1) % 2])
        for _ in range(n):
            xd.copy_(x_host, non_blocking=True)
            yd.copy_(y_host, non_blocking=True)
            last = step(xd, yd).item()
        return last

    e2e_loop(2)
    barrier()
    t0 = time.perf_counter()
    e0.record()
    loss_val = e2e_loop(args.steps)
    e1.record()
    barrier()
    ms_e2e = max_over_ranks(max(e0.elapsed_time(e1), (time.perf_counter() - t0) * 1e3 if world == 1 else 0.0))

    # ---- per-kernel device time of the same step (CUDA events around every launch, separate pass)
    # (every rank runs the steps -- they contain the all-reduce -- but only rank 0 records events)
    prof = None
    nprof = 2
    if rank == 0:
        _lib.profile_begin()
    for _ in range(nprof):
        eager_step(x, y)                                   # eager: events are recorded around every launch
    torch.cuda.synchronize()
    if rank == 0:
        prof = {k: (c / nprof, t / nprof) for k, (c, t) in _lib.profile_end().items()}
    barrier()

    def leave():
        """Multi-rank shutdown in dependency order: the CUDA graph that captured NCCL kernels goes first, then the
        communicator (tearing the communicator down under a live graph hung on the B200 box)."""
        nonlocal graphed
        if graphed is not None:
            graphed.close()
            graphed = None
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            dist.destroy_process_group()

    if rank != 0:
        leave()
        return

    # ---- roofline of the dominant kernel
    trunk = [m for m in model.modules() if isinstance(m, (fnm_b200.SpectralConv2d, fnm_b200.SpectralConv1d))]
    C = trunk[0].in_channels
    if family == "fno1d":
        P1, P2 = 1, xshape[-1] + xshape[-1] // 8
        ny, wb = trunk[0].modes1, trunk[0].weights1.numel() * 8
    elif family == "fnd2d":
        P1, P2 = (s + s // 8 for s in kw["s_outputspace"])
        ny, wb = trunk[0].modes2, 2 * trunk[0].weights1.numel() * 8
    else:
        P1, P2 = xshape[-2] + xshape[-2] // 8, xshape[-1] + xshape[-1] // 8
        ny, wb = trunk[0].modes2, 2 * trunk[0].weights1.numel() * 8
    n_real = sum(p.numel() * (2 if p.is_complex() else 1) for p in model.parameters())
    wl = {"A": 4 * batch * C * P1 * P2, "layers": len(trunk), "ly_frac": 2 * ny / P2, "n_real_params": n_real,
          "w_bytes": wb}
    peak, peak_src = measured_peaks()
    kernels = []
    total_ms = sum(t for _, t in prof.values()) or 1.0
    for name, (cnt, t) in sorted(prof.items(), key=lambda kv: -kv[1][1]):
        ab = kernel_algorithmic_bytes(name, wl, model, cnt)
        kernels.append({"name": name, "launches_per_step": cnt, "ms_per_step": round(t, 4),
                        "share": round(t / total_ms, 4),
                        "algorithmic_GB_per_step": None if ab is None else round(ab / 1e9, 4),
                        "GBps": None if ab is None or t <= 0 else round(ab / 1e6 / t, 1)})
    top = kernels[0]
    # `traffic` is NOT measured by this run: it is the dram read + write bytes per launch of the same kernel from the
    # committed `ncu --set full` capture of this workload (profiles/<file>), named in `traffic_source`
    traffic, traffic_src = None, None
    for tname in ("r02_ncu_traffic.json", "r01_ncu_traffic.json"):
        tpath = os.path.join(ROOT, "profiles", tname)
        if traffic is None and os.path.exists(tpath):
            tj = json.load(open(tpath))
            per = tj.get(args.workload) if isinstance(tj.get(args.workload), dict) else (tj if tj.get("workload") == args.workload else None)
            if per:
                tv = per.get("per_launch_bytes", {}).get(top["name"])
                traffic = tv.get("step_average") if isinstance(tv, dict) else tv
                if traffic is not None:
                    traffic_src = f"committed ncu capture profiles/{tname} (not measured by this run)"
    roofline = {"kernel": top["name"], "bound": "hbm", "achieved": top["GBps"], "peak": peak, "unit": "GB/s",
                "frac": None if top["GBps"] is None else round(top["GBps"] / peak, 4), "traffic": traffic,
                "traffic_source": traffic_src,
                "peak_source": peak_src, "launches_per_step": top["launches_per_step"],
                "ms_per_launch": round(top["ms_per_step"] / max(top["launches_per_step"], 1), 4),
                "algorithmic_bytes_per_step": None if top["algorithmic_GB_per_step"] is None else top["algorithmic_GB_per_step"] * 1e9,
                "trunk_5A_L_GBps": round(5 * wl["A"] * wl["layers"] / 1e6 / (ms / args.steps), 1)}

    # Supplementary tensor-pipe view of the same kernel (in 3xTF32 fp32-parity mode the fused pass is tensor
    # bound, not HBM bound -- DESIGN.md section 4 has the per-role timeline).  Useful FLOPs = 2 * positions * (C + LY) * C per
    # launch; peak = half the measured SUSTAINED dense bf16 rate (TF32), divided by 3 in fp32-parity mode (three TF32
    # products per useful one).  `roofline` above stays the HBM view.
    roofline_tensor = None
    if top["name"] == "pointwise_ysyn" and top["ms_per_step"] > 0:
        bf16 = 1398.6
        mp = os.path.join(ROOT, "MEASURED_PEAKS.json")
        if os.path.exists(mp):
            bf16 = json.load(open(mp)).get("bf16_tflops_sustained", bf16)
        passes = 3 if args.mode == "fp32" else 1
        flops = top["launches_per_step"] * 2.0 * (wl["A"] / (4 * C)) * (C + 2 * ny) * C
        ach = flops / 1e12 / (top["ms_per_step"] / 1e3)
        pk = bf16 / 2 / passes
        roofline_tensor = {"kernel": top["name"], "bound": "tensor", "achieved": round(ach, 1), "peak": round(pk, 1),
                           "unit": "TFLOP/s", "frac": round(ach / pk, 4),
                           "peak_source": f"MEASURED_PEAKS.json bf16_tflops_sustained / 2 (TF32) / {passes} ({args.mode} mode)",
                           "useful_flops_per_step": flops}

    comparator = None
    if world == 1 and not args.no_gpu_comparator:
        # free this arm's step state first: the stock path saves ~10 activation-sized tensors per layer
        if graphed is not None:
            graphed.close()
            graphed = None
        step = eager_step = None
        del grads, opt
        torch.cuda.empty_cache()
        comparator = {tag: gpu_comparator(args.workload, min(args.steps, 5), 3, tf32, dev) for tag, tf32 in (("fp32", False), ("tf32", True))}
    cpu = cpu_baseline(args.workload) if (world == 1 and not args.no_cpu_baseline) else None
    samples = batch * world * args.steps
    line = {
        "metric": "train_samples_per_sec", "value": samples / (ms / 1e3), "unit": "samples/s", "n_gpus": world,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32" if args.mode == "fp32" else "tf32", "data": "synthetic",
        "config": {"workload": f"{args.workload}: {desc}", "batch_per_gpu": batch, "global_batch": batch * world,
                   "parallelism": f"dp{world}", "sm_reserve_backward": sm_reserve, "nccl_max_ctas": os.environ.get("NCCL_MAX_CTAS") if world > 1 else None, "compute_mode": fnm_b200.get_compute_mode(), "cuda_graph": bool(args.graph),
                   "tcgen05": bool(_lib.lib().fnm_has_tcgen05()),
                   "l2": ("working set %.0f MB per activation, larger than the 126 MB L2" % (wl["A"] / 1e6))
                   if wl["A"] > 126e6 else "working set fits L2 (back-to-back steps, no flush)",
                   "optimizer": "complex-aware Adam lr 1e-3 wd 1e-4", "loss": "relative L2, summed over the batch"},
        "clocks": clocks,
        "e2e": {"value": samples / (ms_e2e / 1e3), "unit": "samples/s",
                "h2d_bytes_per_step": (x_host.numel() + y_host.numel()) * 4, "d2h_bytes_per_step": 4,
                "ms_per_step": ms_e2e / args.steps, "last_loss": loss_val,
                "input_pipeline": "step i + 1's host-to-device copy overlaps step i (double buffered); every step's loss is copied to "
                                  "pinned host memory and read by the host one step later, while the next step runs"
                                  if graphed is not None else "copy, step, loss read-back in sequence"},
        "gpu_launches": int(launches),
        "roofline": roofline,
        "roofline_tensor": roofline_tensor,
        "kernels": kernels,
        "cpu_baseline": cpu,
        "gpu_comparator": comparator,
    }
    # teardown first, the JSON line LAST: with NCCL_DEBUG=INFO the communicator's teardown lines of every rank also go to
    # stdout, and the one JSON line must stay the last line of the output
    leave()
    if world > 1:
        time.sleep(0.5)
    print(json.dumps(line))
    sys.stdout.flush()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference", "stock-gpu"])
    ap.add_argument("--workload", default="cfg5", choices=sorted(WORKLOADS))
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-gpu-comparator", action="store_true",
                    help="skip the stock-PyTorch GPU comparator (cuFFT / cuBLAS / cuDNN path of the reference) at N = 1")
    ap.add_argument("--mode", default="fp32", choices=["fp32", "tf32"],
                    help="fp32: 3xTF32 parity mode (rel-L2 <= 1e-5, the headline); tf32: single tensor-core pass (<= 5e-3)")
    ap.add_argument("--graph", dest="graph", action="store_true", default=True,
                    help="replay the step from one CUDA graph (default)")
    ap.add_argument("--no-graph", dest="graph", action="store_false")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", 0))
    world = int(os.environ.get("WORLD_SIZE", 1))
    if args.impl == "reference":
        run_reference(args, rank)
        return
    if args.impl == "stock-gpu":
        run_stock_gpu(args, rank)
        return
    if args.warmup < 3:
        args.warmup = 3
    if world != args.gpus and world == 1 and args.gpus > 1:
        # launched without torchrun: re-exec under torch.distributed.run
        cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={args.gpus}",
               "--master-addr", "127.0.0.1", "--master-port", "29511", os.path.abspath(__file__)] + sys.argv[1:]
        sys.exit(subprocess.call(cmd))
    run_ours(args, rank, world)


if __name__ == "__main__":
    main()
