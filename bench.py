#!/usr/bin/env python
"""bench.py — headline benchmark of the per-block audio-graph DSP path (BASELINE.json metric:
mono-equivalent samples/s through the graph).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference] [--workload c2]

A *step* is one pass of the hot path over one batch of synthetic input:
  c2 (default, BASELINE configs[1]): 1024 stereo voices per GPU, gain -> pan -> master-bus sum,
      256-frame blocks, 256 consecutive blocks per step (65536 frames; 512 MiB of f32 input per GPU,
      larger than the 126 MB L2, so every step streams from HBM).
N > 1 is launched by torchrun (one rank per GPU); voices shard by rank (weak scaling).

Prints ONE JSON line on rank 0. `value` = device-timed, inputs resident in HBM. `e2e` = same metric through
the host-buffer C-ABI call (pinned host memory; H2D + D2H inside the timed region).
`--impl reference` times the CPU oracle (the C++ restatement of the reference's Rust path — Rust cannot be
built in this image) on all host cores over a bounded sample of the same workload.
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time
from concurrent.futures import ThreadPoolExecutor
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "oracle"))

F32 = np.float32
SR = 48000
WORKLOADS = {
    # name: voices per GPU, channels, block frames, blocks per step
    "c2": dict(voices=1024, ch=2, block=256, blocks=256, bus=True, bytes_per_sample=4.004, kernel_class=1,
               kernel="chain_kernel<VEC=4,CIN=2,BUS,4 voices/warp,16 warps> (gain->pan->bus tree)",
               desc="c2: 1024 stereo voices/GPU, gain->pan->master-bus sum, 256-frame blocks, 256 blocks/step"),
    "c3": dict(voices=4096, ch=2, block=512, blocks=32, bus=False, bytes_per_sample=16.0, kernel_class=3,
               kernel="biquad_delay_lanes<NS=4,L=4,DELAY,FULL> (4-stage biquad cascade + 12000-frame delay ring, stage-parallel lanes)",
               desc="c3: 4096 stereo voices/GPU, 4-stage biquad cascade + 12000-frame delay line, 512-frame blocks, 32 blocks/step"),
    "c4": dict(voices=256, ch=2, block=512, blocks=16, bus=False, bytes_per_sample=8.0, kernel_class=3, ir_len=48000,
               kernel="reverb_gemm_kernel (tcgen05.mma kind::f16 M128 N256 K16, TMEM accumulators, TMA 128B-swizzle operands)",
               desc="c4: 256 stereo voices/GPU, FIR convolutional reverb, 48000-tap stereo IR, bf16 tcgen05, 512-frame blocks, 16 blocks/step"),
    "c5": dict(voices=8192, ch=2, block=512, blocks=2, bus=True, bytes_per_sample=8.0, kernel_class=3, ir_len=48000,
               kernel="reverb_gemm_kernel (tcgen05) + biquad_delay_lanes; chain_kernel before and after",
               desc="c5: 8192 stereo voices/GPU (65536 over 8), gain->pan->4-stage biquad->48000-tap FIR reverb->master-bus sum, 512-frame blocks, 2 blocks/step"),
    # config 5 with its real source: every voice is a looping SamplerNode reading an f32 stereo sample resource in HBM
    "c5s": dict(voices=8192, ch=2, block=512, blocks=2, bus=True, bytes_per_sample=8.0, kernel_class=3, ir_len=48000, src="sampler",
                kernel="reverb_gemm_kernel (tcgen05) + biquad_delay_lanes; sampler_kernel source, chain_kernel before and after",
                desc="c5s: 8192 looping stereo sampler voices/GPU (256 two-second f32 resources, 188 MiB in HBM) -> gain->pan->4-stage biquad->48000-tap FIR reverb->master-bus sum, 512-frame blocks, 2 blocks/step"),
}
REVERB_WORKLOADS = ("c4", "c5", "c5s")


def synth(shape, seed):
    rng = np.random.Generator(np.random.PCG64(seed))
    r = rng.integers(0, 1 << 24, size=shape, dtype=np.uint32)
    return (r.astype(F32) * F32(2.0 ** -23) - F32(1.0)).astype(F32)


def voice_params(V, seed):
    rng = np.random.default_rng(seed)
    return (25 + 75 * rng.random(V)).astype(F32), rng.uniform(-1, 1, V).astype(F32)  # never muted


def biquad_params(fw, lib, V, seed):
    """RBJ low-pass / peaking, fc log-uniform 200 Hz - 8 kHz, Q in [0.5, 2] at 48 kHz."""
    rng = np.random.default_rng(seed)
    k = np.zeros((V, 4, 5), F32)
    for v in range(V):
        for s in range(4):
            k[v, s] = fw.design_rbj(lib, 0 if s % 2 == 0 else 4, 200.0 * 40.0 ** rng.random(), rng.uniform(0.5, 2.0), rng.uniform(-3, 3), SR)
    return k


def build_graph(fw, lib, workload, V, block, device, seed):
    """The voice graph of a workload, on `lib` (the CUDA product, or the CPU oracle for the baseline legs)."""
    w = WORKLOADS[workload]
    sampler_src = w.get("src") == "sampler"
    cx = fw.FirewheelGraphCtx(lib, fw.AudioGraphConfig(num_graph_inputs=0 if sampler_src else 2, num_graph_outputs=2, num_voices=V, master_bus=w["bus"], device=device))
    g = cx.graph
    if workload == "c2":
        pct, pan = voice_params(V, seed)
        nodes = [g.add_node(2, 2, fw.VolumeNode(100.0)), g.add_node(2, 2, fw.PanNode(0.0))]
        g.set_percent_volume(nodes[0], pct)
        g.set_pan(nodes[1], pan)
    elif workload in REVERB_WORKLOADS:
        L = w["ir_len"]
        rng = np.random.default_rng(0x1200)
        ir = rng.standard_normal((2, L)) * np.exp(-6.9 * np.arange(L) / L)
        ir = (ir / np.sqrt((ir ** 2).sum(axis=1, keepdims=True))).astype(F32)
        nodes = [g.add_node(2, 2, fw.ConvReverbNode(ir))]
        if workload in ("c5", "c5s"):
            pct, pan = voice_params(V, seed)
            pre = [g.add_node(2, 2, fw.VolumeNode(100.0)), g.add_node(2, 2, fw.PanNode(0.0)), g.add_node(2, 2, fw.BiquadNode(4))]
            g.set_percent_volume(pre[0], pct); g.set_pan(pre[1], pan); g.set_biquad_coeffs(pre[2], biquad_params(fw, lib, V, seed))
            nodes = pre + nodes
    else:
        nodes = [g.add_node(2, 2, fw.BiquadNode(4)), g.add_node(2, 2, fw.DelayNode(12000))]
        g.set_biquad_coeffs(nodes[0], biquad_params(fw, lib, V, seed))
    smp = None
    if sampler_src:
        smp = g.add_node(0, 2, fw.SamplerNode(100.0))
        prev = smp
    else:
        prev = g.graph_in_node()
    for n in nodes + [g.graph_out_node()]:
        for c in range(2):
            g.connect(prev, c, n, c, False)
        prev = n
    proc = cx.activate(SR, 0 if sampler_src else 2, 2, block)
    if proc is None:
        raise RuntimeError("activate failed")
    st = cx.update()
    if st.graph_error is not None:
        raise RuntimeError(f"compile failed: {st.graph_error} {cx.last_error()}")
    if sampler_src:
        # 256 two-second stereo f32 resources (188 MiB: larger than L2); voice v loops resource v % n_res from its own offset
        n_res = min(256, V)
        handles = [g.create_sample_resource(synth((2, 2 * SR), 0x5A000000 + seed * 4096 + r)) for r in range(n_res)]
        for v in range(V):
            g.sampler_set_sample(smp, handles[v % n_res], True, voice=v)
            g.sampler_set_loop_range(smp, "full", voice=v)
            g.sampler_set_playhead(smp, ((v // n_res) * 2731 % (2 * SR)) / SR, voice=v)
        g.sampler_play(smp)
    return cx, proc


# ---- clocks (B200_PROFILING.md: sample DURING the timed region) ---------------------------------
class ClockSampler:
    Q = "index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, gpu_index):
        self.idx, self.rows, self.stop_evt, self.th = gpu_index, [], threading.Event(), None

    def _once(self):
        try:
            out = subprocess.run(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-i", str(self.idx)],
                                 capture_output=True, text=True, timeout=5).stdout.strip()
            if out:
                self.rows.append([c.strip() for c in out.split(",")])
                return
        except Exception:
            pass
        try:
            import pynvml
            pynvml.nvmlInit()
            h = pynvml.nvmlDeviceGetHandleByIndex(self.idx)
            sm = pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM)
            mx = pynvml.nvmlDeviceGetMaxClockInfo(h, pynvml.NVML_CLOCK_SM)
            r = pynvml.nvmlDeviceGetCurrentClocksEventReasons(h) if hasattr(pynvml, "nvmlDeviceGetCurrentClocksEventReasons") else pynvml.nvmlDeviceGetCurrentClocksThrottleReasons(h)
            f = lambda bit: "Active" if r & bit else "Not Active"
            self.rows.append([str(self.idx), str(sm), str(mx), "0", f(0x8), f(0x40), f(0x20), f(0x4)])
        except Exception:
            pass

    def start(self):
        def loop():
            while not self.stop_evt.is_set():
                self._once()
                self.stop_evt.wait(0.2)
        self.th = threading.Thread(target=loop, daemon=True)
        self.th.start()

    def stop(self):
        self.stop_evt.set()
        if self.th:
            self.th.join(timeout=10)
        sm = [float(r[1]) for r in self.rows if r[1].replace(".", "").isdigit()]
        mx = [float(r[2]) for r in self.rows if r[2].replace(".", "").isdigit()]
        reasons = set()
        for r in self.rows:
            for name, col in (("hw_slowdown", 4), ("hw_thermal_slowdown", 5), ("sw_thermal_slowdown", 6), ("sw_power_cap", 7)):
                if len(r) > col and r[col] == "Active":
                    reasons.add(name)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(self.rows)}


def peaks():
    p = ROOT / "MEASURED_PEAKS.json"
    if p.exists():
        d = json.loads(p.read_text())
        return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


# ---- CPU oracle legs ---------------------------------------------------------------------------------
def oracle_rate(V, block, n_blocks, threads, seed=7, steps=1, warmup=0, workload="c2"):
    """Mono-equivalent samples/s of the CPU oracle on V voices x n_blocks blocks, voices split over `threads`
    replicas (each a disjoint voice range with its own partial bus; the reference itself is single-threaded)."""
    import firewheel_b200 as fw
    import pyoracle
    lib = pyoracle.load()
    T = block * n_blocks
    bus = WORKLOADS[workload]["bus"]
    bounds = np.linspace(0, V, threads + 1).astype(int)
    parts = []
    for i in range(threads):
        lo, hi = bounds[i], bounds[i + 1]
        if hi <= lo:
            continue
        cx, proc = build_graph(fw, lib, workload, hi - lo, block, 0, seed * 100 + i)
        x = synth((hi - lo, 0 if WORKLOADS[workload].get("src") == "sampler" else 2, T), seed * 1000 + i)
        out = np.zeros((2, T) if bus else (hi - lo, 2, T), F32)
        parts.append((cx, proc, x, out))

    def run(p):
        cx, proc, x, out = p
        rc, _ = proc.process_planar(x, out, x.shape[1], 2, T)
        assert rc == 0

    pool = ThreadPoolExecutor(len(parts)) if len(parts) > 1 else None  # persistent: thread start-up is not the reference's cost

    def one_step():
        if pool is None:
            run(parts[0])
        else:
            list(pool.map(run, parts))
            if bus:
                mix = parts[0][3].copy()
                for p in parts[1:]:
                    mix += p[3]  # top of the mix tree over replicas
    for _ in range(warmup):
        one_step()
    t0 = time.perf_counter()
    for _ in range(steps):
        one_step()
    dt = time.perf_counter() - t0
    if pool is not None:
        pool.shutdown()
    for cx, proc, _, _ in parts:
        proc.free(); cx.update(); cx.free()
    return V * 2 * T * steps / dt, dt / steps


def run_reference(args, rank, world):
    if rank != 0:
        return
    w = WORKLOADS[args.workload]
    cores = os.cpu_count() or 1
    V = w["voices"] * max(args.gpus, 1)
    n_blocks = w["blocks"]  # c2 / c3: the full step (a few ms per replica thread), so dispatch overhead does not flatter the GPU
    if args.workload in REVERB_WORKLOADS:
        V, n_blocks = max(cores, 2) * 1, 1  # direct-form FIR on the CPU: 96 kflop per output sample
    val, sec_per_step = oracle_rate(V, w["block"], n_blocks, cores, steps=args.steps, warmup=args.warmup, workload=args.workload)
    sample = f"{V} voices x {n_blocks} of {w['blocks']} blocks per step, {cores} replica threads over disjoint voice ranges"
    line = {"impl": "reference", "metric": "mono_equiv_samples_per_sec", "value": val, "unit": "samples/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": sec_per_step * 1e3, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": w["desc"], "voices_total": V, "block_frames": w["block"], "blocks_per_step": n_blocks,
                       "note": "CPU oracle = C++ restatement of the reference's Rust path (no Rust toolchain in this image)"},
            "cpu_baseline": {"value": val, "unit": "samples/s", "cores": cores, "kind": "port", "sample": sample},
            "e2e": {"value": val, "unit": "samples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


# ---- the product arm ---------------------------------------------------------------------------------
def run_b200(args, rank, world, local_rank):
    import firewheel_b200 as fw
    lib = fw.load()  # raises if the CUDA library is missing: no CPU fallback
    if lib.device_count() <= local_rank:
        raise RuntimeError("no CUDA device for this rank: " + lib.last_device_error().decode())
    dist = None
    if world > 1:
        # NCCL prints its version banner on stdout at NCCL_DEBUG >= VERSION; keep stdout for the single JSON line
        os.environ["NCCL_DEBUG_FILE"] = "/dev/stderr"
        if os.environ.get("NCCL_DEBUG", "").upper() in ("VERSION", "WARN"):
            os.environ["NCCL_DEBUG"] = "NONE"
        import torch
        import torch.distributed as dist_mod
        torch.cuda.set_device(local_rank)
        dist_mod.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
        dist = dist_mod

    w = WORKLOADS[args.workload]
    V, C, F, KB = w["voices"], w["ch"], w["block"], w["blocks"]
    T = F * KB
    cx, proc = build_graph(fw, lib, args.workload, V, F, local_rank, 1000 + rank)
    if dist and w["bus"]:
        # the master bus crosses ranks inside the product: NCCL all-gather of the per-rank buses over NVLink, then the
        # top levels of the balanced tree in rank order (bit-identical on every rank)
        import ctypes as _ct
        ids = [bytes(128)]
        if rank == 0:
            buf = (_ct.c_uint8 * 128)()
            if lib.comm_unique_id(buf) != 0:
                raise RuntimeError("ncclGetUniqueId failed: " + lib.last_device_error().decode())
            ids = [bytes(buf)]
        dist.broadcast_object_list(ids, src=0)
        if proc.comm_init(rank, world, ids[0]) != 0:
            raise RuntimeError("ncclCommInitRank failed: " + lib.last_device_error().decode())
    Cin = 0 if w.get("src") == "sampler" else C  # sampler voices read their sample resources from HBM, not a stream input
    in_bytes = V * Cin * T * 4
    out_bytes = C * T * 4 if w["bus"] else V * C * T * 4

    # synthetic input straight into pinned host memory, then resident in HBM
    h_in = lib.host_alloc_pinned(in_bytes) if in_bytes else 0
    h_out = lib.host_alloc_pinned(out_bytes)
    if (in_bytes and not h_in) or not h_out:
        raise RuntimeError("pinned allocation failed")
    import ctypes
    if in_bytes:
        x = np.ctypeslib.as_array(ctypes.cast(h_in, ctypes.POINTER(ctypes.c_float)), shape=(V, C, T))
        chunk = 64
        for v0 in range(0, V, chunk):
            x[v0:v0 + chunk] = synth((min(chunk, V - v0), C, T), 0xF17E0000 + rank * 65536 + v0)
    y = np.ctypeslib.as_array(ctypes.cast(h_out, ctypes.POINTER(ctypes.c_float)), shape=(C, T) if w["bus"] else (V, C, T))
    d_in, d_out = (lib.dev_malloc(local_rank, in_bytes) if in_bytes else 0), lib.dev_malloc(local_rank, out_bytes)
    if (in_bytes and not d_in) or not d_out:
        raise RuntimeError("device allocation failed: " + lib.last_device_error().decode())
    if in_bytes:
        proc.h2d(d_in, h_in, in_bytes)
    proc.sync()

    def barrier():
        proc.sync()
        if dist:
            dist.barrier()

    def max_over_ranks(v):
        if not dist:
            return v
        import torch
        t = torch.tensor([v], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    # ---- device-timed pass (inputs resident in HBM) ----
    for _ in range(max(args.warmup, 3)):
        assert proc.process_planar_device(d_in, d_out, Cin, C, T) == 0
    barrier()
    clocks = ClockSampler(local_rank)
    clocks.start()
    launches0 = proc.kernel_launches()
    proc.event_record(0)
    for _ in range(args.steps):
        assert proc.process_planar_device(d_in, d_out, Cin, C, T) == 0
    proc.event_record(1)
    barrier()
    ms_total = max_over_ranks(proc.event_elapsed_ms(0, 1))
    launches = proc.kernel_launches() - launches0
    # second pass, same steps, with a CUDA-event pair around every kernel class (these events serialise the
    # programmatic-dependent-launch overlap, so they stay out of the headline pass)
    proc.profile(True)
    for _ in range(args.steps):
        assert proc.process_planar_device(d_in, d_out, Cin, C, T) == 0
    prof_ms, prof_n = proc.profile_read()
    proc.profile(False)
    clk = clocks.stop()
    ms_per_step = ms_total / args.steps
    samples_per_step = V * C * T * world
    value = samples_per_step / (ms_per_step * 1e-3)

    # parity spot check of the timed configuration is in tests/; here only a finiteness guard on the result
    proc.d2h(h_out, d_out, out_bytes)
    proc.sync()
    assert np.all(np.isfinite(y[..., ::97])) and float(np.abs(y[..., ::97]).max()) > 0.0

    # ---- end-to-end pass: host buffers through the C-ABI call, H2D + D2H inside the timed region ----
    e2e_steps = min(args.steps, 10)
    for _ in range(2):
        rc, _ = proc.process_planar(h_in, h_out, Cin, C, T)
        assert rc == 0
    barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        rc, _ = proc.process_planar(h_in, h_out, Cin, C, T)
        assert rc == 0
    proc.sync()
    e2e_s = max_over_ranks((time.perf_counter() - t0) / e2e_steps)
    e2e_value = samples_per_step / e2e_s

    # ---- roofline of the dominant kernel (fused chain + bus), CUDA events on the launching stream ----
    peak, peak_src = peaks()
    kc = w["kernel_class"]
    chain_ms = prof_ms[kc] / max(args.steps, 1)  # all launches of the dominant kernel class in one step
    # c2 reads V*C*T f32 and writes the C*T bus; c3 moves in + out + delay-ring read + write = 16 B/sample
    algo_bytes = 4 * C * T * (V + 1) if args.workload == "c2" else int(w["bytes_per_sample"] * V * C * T)
    achieved = algo_bytes / (chain_ms * 1e-3) / 1e9 if chain_ms > 0 else 0.0
    traffic = None
    tp = ROOT / "profiles" / ("r01_chain_traffic.json" if args.workload == "c2" else f"r01_{args.workload}_traffic.json")
    if tp.exists():
        try:
            traffic = json.loads(tp.read_text()).get("dram_bytes_per_launch")
        except Exception:
            traffic = None
    if args.workload in REVERB_WORKLOADS:  # tensor-pipe roofline: dense direct-form count 2*L flop per output sample
        pk = ROOT / "MEASURED_PEAKS.json"
        peak, peak_src = (float(json.loads(pk.read_text())["bf16_tflops"]), "measured (MEASURED_PEAKS.json bf16_tflops, burst)") if pk.exists() else (1590.0, "fallback (B200_PROFILING.md 1.59 PFLOP/s)")
        flops = 2.0 * w["ir_len"] * V * C * T
        achieved = flops / (chain_ms * 1e-3) / 1e12 if chain_ms > 0 else 0.0
    roofline = {"bound": "hbm" if args.workload not in REVERB_WORKLOADS else "tensor", "kernel": w["kernel"], "achieved": achieved, "peak": peak, "unit": "GB/s" if args.workload not in REVERB_WORKLOADS else "TFLOP/s",
                "frac": achieved / peak, "traffic": traffic, "peak_source": peak_src, "kernel_ms": chain_ms,
                "algorithmic_bytes_per_launch": algo_bytes, "algorithmic_flops_per_launch": (2.0 * w["ir_len"] * V * C * T) if args.workload in REVERB_WORKLOADS else None,
                "step_share": {"control_ms": prof_ms[0] / max(prof_n[0], 1), "chain_ms": prof_ms[1] / max(prof_n[1], 1),
                               "combine_ms": prof_ms[2] / max(prof_n[2], 1), "temporal_ms": prof_ms[3] / max(prof_n[3], 1)}}

    cpu = None
    if rank == 0 and not os.environ.get("FW_BENCH_SKIP_CPU"):  # (experiments on multi-GPU boxes skip the 10 s CPU leg)
        n_blocks = 64
        n_blocks = 64 if args.workload == "c2" else 2
        Vc = V if args.workload not in REVERB_WORKLOADS else 2  # the direct-form FIR oracle needs ~0.2 s per voice-block
        if args.workload in REVERB_WORKLOADS:
            n_blocks = 1
        rate, sec = oracle_rate(Vc, F, n_blocks, 1, steps=1, warmup=0, workload=args.workload)
        if sec < 2.0 and args.workload not in REVERB_WORKLOADS:  # size the sample towards ~10 s of CPU work
            n_blocks = int(min(KB * 8, max(n_blocks, n_blocks * 10.0 / max(sec, 1e-3))))
            rate, sec = oracle_rate(Vc, F, n_blocks, 1, steps=1, warmup=0, workload=args.workload)
        elif args.workload in REVERB_WORKLOADS and sec < 5.0:
            Vc = int(min(64, max(2, Vc * 10.0 / max(sec, 1e-3))))
            rate, sec = oracle_rate(Vc, F, n_blocks, 1, steps=1, warmup=0, workload=args.workload)
        cpu = {"value": rate, "unit": "samples/s", "cores": 1, "kind": "port",
               "sample": f"{Vc} voices x {n_blocks} blocks of {F} frames, 1 thread (the reference's execution model), {sec:.1f} s"}

    if rank == 0:
        line = {"metric": "mono_equiv_samples_per_sec", "value": value, "unit": "samples/s", "n_gpus": world, "steps": args.steps,
                "warmup": max(args.warmup, 3), "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                "dtype": "f32", "data": "synthetic",
                "config": {"workload": w["desc"], "voices_per_gpu": V, "channels": C, "block_frames": F, "blocks_per_step": KB,
                           "l2": (f"inputs larger than L2 ({in_bytes >> 20} MiB per GPU per step)" if in_bytes else "sample pool larger than L2 (188 MiB per GPU), every voice at its own offset"), "layout": "planar [voice][ch][frame]",
                           "parallelism": f"voices sharded over {world} rank(s)" + (("; master bus = peer-memory (NVLink) push + rank-ordered tree" if os.environ.get("FW_EXCHANGE") == "p2p" else "; master bus = NCCL all-gather on a high-priority side stream + rank-ordered tree") if (world > 1 and w["bus"]) else "")},
                "clocks": clk,
                "e2e": {"value": e2e_value, "unit": "samples/s", "h2d_bytes_per_step": in_bytes * world, "d2h_bytes_per_step": out_bytes * world,
                        "steps": e2e_steps, "ms_per_step": e2e_s * 1e3, "api": "fw_processor_process_planar (pinned host buffers)"},
                "gpu_launches": int(launches),
                "roofline": roofline, "cpu_baseline": cpu}
        print(json.dumps(line), flush=True)
    proc.free()
    cx.update()
    cx.free()
    if dist:
        dist.barrier()
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="c2", choices=sorted(WORKLOADS))  # c2 is the headline (BASELINE configs[1])
    args = ap.parse_args()
    rank, world = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank, world)
    else:
        run_b200(args, rank, world, local_rank)


if __name__ == "__main__":
    main()
