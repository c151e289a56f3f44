#!/usr/bin/env python
"""bench.py — train rays/s (fwd+bwd) of the tensorf-jax per-ray hot path on N B200s.

    python bench.py --gpus 1 --steps 20 --warmup 5
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 \
        --master-port P bench.py --gpus N --steps K --warmup W
    python bench.py --impl reference ...      # CPU arm: the oracle restatement on the host cores

A "step" is one pass of the hot path (render_rays forward + loss + reverse w.r.t. every leaf of
LearnableParams; training.py:108-156) over one batch of synthetic rays.  Workload at every N:
BASELINE.json configs[1] per GPU (lego shape, R=4096 rays x N=256 density samples, K=38
appearance samples, 128^3 grid); for N>1 the rays are sharded (weak scaling: 4096 rays per
GPU), the loss is the mean over the GLOBAL batch and the gradients are sum-allreduced with
NCCL every step.  Adam/LR (training.py:158-204) is outside the measured path.
Rank 0 prints ONE JSON line.
"""
from __future__ import annotations

import argparse
import json
import os
import pathlib
import subprocess
import sys
import threading
import time

ROOT = pathlib.Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT / "tensorf-jax_b200"))

import numpy as np  # noqa: E402
import torch  # noqa: E402

from tensorf_b200 import synthetic as S  # noqa: E402

METRIC = "train_rays_per_sec_fwd_bwd"
UNIT = "rays/s"


def workload_from_name(name: str) -> S.Workload:
    if name == "lego_256":  # BASELINE.json configs[1] as named ("4096 rays x 256 samples")
        return S.lego_workload(R=4096, G=128, N=256, K=38, name="lego_G128_R4096_N256_K38 (BASELINE configs[1])")
    if name == "lego_221":  # what the reference would execute at G=128 (training.py:115-118)
        return S.lego_workload(R=4096, G=128)
    if name == "lego_300":  # configs[2] per-GPU shard at 8 GPUs (16384 global rays)
        return S.lego_workload(R=2048, G=300)
    if name == "lego_300_4096":
        return S.lego_workload(R=4096, G=300)
    if name == "dozer_300":
        return S.dozer_workload(R=2048, G=300)
    if name == "dozer_128":
        return S.dozer_workload(R=2048, G=128)
    raise SystemExit(f"unknown workload {name}")


def measured_peaks():
    """(HBM GB/s, dense bf16 TFLOP/s sustained, where they come from)."""
    p = ROOT / "MEASURED_PEAKS.json"
    if p.exists():
        d = json.loads(p.read_text())
        return (float(d["hbm_gbs"]), float(d.get("bf16_tflops_sustained", 1400.0)),
                "measured (MEASURED_PEAKS.json: hbm_gbs burst copy; bf16_tflops_sustained for kernels timed inside the step)")
    return 6650.0, 1400.0, "fallback (B200_PROFILING.md: 6.65 TB/s, ~1.4 PFLOP/s sustained bf16)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region."""

    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.rows, self.proc = [], None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "100", "-i", str(index)], stdout=subprocess.PIPE, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        self.thread.join(timeout=2)
        sm = [float(r[0]) for r in self.rows if len(r) >= 6 and r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in self.rows if len(r) >= 6 and r[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({n for r in self.rows if len(r) >= 6 for n, v in zip(names, r[2:6]) if v.lower().startswith("active")})
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": reasons,
                "samples": len(sm)}


# ---------------------------------------------------------------------------------------------
# CPU arm: the oracle (PyTorch-CPU fp32 restatement of the reference; JAX is not in the image)
# ---------------------------------------------------------------------------------------------
def cpu_reference_rays_per_s(w: S.Workload, sample_rays: int, steps: int, warmup: int):
    sys.path.insert(0, str(ROOT / "oracle"))
    sys.path.insert(0, str(ROOT / "tests"))
    import tensorf_oracle as O
    from helpers import oracle_cfgs, oracle_inputs

    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    inp = S.make_inputs(w, R=sample_rays)
    ws = S.Workload(**{**w.__dict__, "R": sample_rays})
    cfg, mc = oracle_cfgs(ws)
    oi = oracle_inputs(inp, torch.float32)

    def step():
        O.loss_and_grads(cfg, mc, oi["params"], w.contracted, oi["aabb"], oi["origins"], oi["directions"],
                         oi["camera_indices"], oi["colors"], oi["jitter"], oi["gumbel"])

    for _ in range(warmup):
        step()
    t0 = time.perf_counter()
    for _ in range(steps):
        step()
    dt = (time.perf_counter() - t0) / steps
    return sample_rays / dt, dt * 1e3, cores


def run_reference_arm(args, w):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    sample = 512
    steps = max(1, min(args.steps, 10))
    warmup = max(1, min(args.warmup, 2))
    v, ms, cores = cpu_reference_rays_per_s(w, sample, steps, warmup)
    desc = f"{sample}-ray slice of the {w.name} batch per step, fwd+bwd via torch.autograd, fp32, {cores} threads"
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": steps,
        "warmup": warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": {"workload": w.name, "R_per_step": sample, "N": w.N, "K": w.K, "G": w.G, "cd": w.cd, "ca": w.ca,
                   "note": "CPU restatement of the reference (oracle/tensorf_oracle.py); JAX is not installable in this "
                           "image so the reference's own JAX-CPU path cannot run"},
        "cpu_baseline": {"value": v, "unit": UNIT, "cores": cores, "kind": "port", "sample": desc},
        "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }))


# ---------------------------------------------------------------------------------------------
# B200 arm
# ---------------------------------------------------------------------------------------------
def bench_render(ops, dev, world, rank, dist, frames=2, warm=1):
    """render_360 as a job (render_360.py:118-161 through render.py:49-102, BASELINE configs[4] shapes): full 800x800
    frames of the 300^3 lego model at N=512 / K=128, chunks of 16384 rays with the ragged last chunk, per mode
    (RGB, median distance, mean distance).  Per frame and rank: `tensorf_pixel_rays` on the device -> chunk loop ->
    ONE device-to-host copy of the rank's rows into pinned memory, all inside the timed region.  The frame is dealt to
    the ranks in interleaved 16-row stripes (`dist.stripe_rows`), no collective.  Whole-job Mpix/s = frame pixels x
    frames / max-over-ranks device time (CUDA events around the frames, including the copies)."""
    from tensorf_b200 import networks, render as trender
    w = S.render360_workload()
    H = W_ = 800
    flat = {k: torch.from_numpy(v).to(dev) for k, v in S.make_params(w.G, w.cd, w.ca, w.feat_freqs, w.view_freqs, None, 0).items()}
    lp = trender.LearnableParams.from_flat(flat, scene_contraction=False)
    mlp = networks.FeatureMlp(feature_n_freqs=w.feat_freqs, viewdir_n_freqs=w.view_freqs)
    aabb = torch.from_numpy(w.aabb()).to(dev)
    jitter, gumbel = S.make_noise(w.N, w.R, False, 2)
    noise = {"jitter": torch.from_numpy(jitter).to(dev), "gumbel": torch.from_numpy(gumbel).to(dev)}
    cams = [S.frame_camera(W_, H, angle=0.3 + 2 * 3.14159265 * i / 120) for i in range(frames + warm)]  # consecutive poses of the orbit
    out = {"workload": w.name, "frame": [H, W_], "rays_per_chunk": w.R, "frames_timed": frames,
           "tiling": f"{__import__('tensorf_b200.dist', fromlist=['x']).balanced_stripe(H, world)}-row stripes round-robin over {world} rank(s) (every rank the same number of rows), no collective",
           "e2e": "pixel rays generated on the device, one D2H copy of the rank's rows per frame into pinned memory (inside the timed region)"}
    for label, mode in (("rgb", trender.RenderMode.RGB), ("dist_median", trender.RenderMode.DIST_MEDIAN), ("dist_mean", trender.RenderMode.DIST_MEAN)):
        cfg = trender.RenderConfig(near=0.1, far=10.0, mode=mode, density_samples_per_ray=w.N, appearance_samples_per_ray=w.K)
        fr = trender.FrameRenderer(mlp, lp, aabb, cfg, H, W_, batch_size=w.R, rank=rank, world=world)
        ns = {k: v for k, v in noise.items() if k == "jitter" or mode is trender.RenderMode.RGB}
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        for i, cam in enumerate(cams):
            if i == warm:
                if dist is not None:
                    dist.barrier()
                torch.cuda.synchronize()
                ev0.record()
            fr.render(cam, i, ns, sync=False)
        ev1.record()
        torch.cuda.synchronize()
        t = torch.tensor([ev0.elapsed_time(ev1)], dtype=torch.float64, device=dev)
        if dist is not None:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item()) / frames
        bytes_d2h = fr.frame_host.numel() * 4
        out[label] = {"mpix_per_s": H * W_ / (ms * 1e-3) / 1e6, "ms_per_frame": ms, "rays_this_rank": fr.n,
                      "chunks_per_frame_this_rank": (fr.n + w.R - 1) // w.R, "ragged_last_chunk": fr.n % w.R, "d2h_bytes_per_frame_this_rank": bytes_d2h,
                      "finite": bool(torch.isfinite(fr.frame_host).all())}
        del fr
        torch.cuda.empty_cache()
    return out


def bench_optimizer(ops, dev, params, grads, peak_gbs, reps=10):
    """The two callers right after the reverse pass, timed on their own (NOT part of the
    headline step): one `tensorf_adam_step` over every leaf (training.py:183-201; 16 B read + 12 B written per
    parameter) and one `tensorf_vm_resize` of both factor sets to the next grid of the lego schedule
    (training.py:245-276; bytes = read old + write new)."""
    names = list(params.keys())
    p = [params[k].clone() for k in names]
    mu = [torch.zeros_like(x) for x in p]
    nu = [torch.zeros_like(x) for x in p]
    g = [grads[k] for k in names]
    call = ops.AdamCall(p, mu, nu, [-(0.02 if k.startswith(("density_", "appearance_")) else 1e-3) for k in names])
    n = sum(x.numel() for x in p)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    for i in range(3):
        call.step(g, count=i)
    torch.cuda.synchronize()
    # the kernel is ~15 us at 128^3: replay `reps` captured steps so host launch overhead is not what is timed
    graph, side = torch.cuda.CUDAGraph(), torch.cuda.Stream()
    with torch.cuda.stream(side):
        with torch.cuda.graph(graph, stream=side):
            for i in range(reps):
                call.step(g, count=3 + i)
    graph.replay()
    torch.cuda.synchronize()
    e0.record()
    graph.replay()
    e1.record()
    torch.cuda.synchronize()
    adam_ms = e0.elapsed_time(e1) / reps
    out = {"adam": {"ms": adam_ms, "parameters": n, "GB/s": 28 * n / (adam_ms * 1e-3) / 1e9,
                    "frac_of_hbm_peak": 28 * n / (adam_ms * 1e-3) / 1e9 / peak_gbs, "launches_per_step": 2}}
    # the same kernel where it is truly DRAM-bound: the final 300^3 grid (17.3 M parameters, 485 MB per step > L2)
    shapes300 = {"density_vector": (3, 16, 300), "density_matrix": (3, 16, 300, 300), "appearance_vector": (3, 48, 300),
                 "appearance_matrix": (3, 48, 300, 300)}
    p3 = [torch.randn(s_, device=dev) * 0.1 for s_ in shapes300.values()]
    g3 = [torch.randn(s_, device=dev) * 1e-3 for s_ in shapes300.values()]
    call3 = ops.AdamCall(p3, [torch.zeros_like(x) for x in p3], [torch.zeros_like(x) for x in p3], [-0.02] * 4)
    n3 = sum(x.numel() for x in p3)
    for i in range(3):
        call3.step(g3, count=i)
    torch.cuda.synchronize()
    e0.record()
    for i in range(reps):
        call3.step(g3, count=3 + i)
    e1.record()
    torch.cuda.synchronize()
    ms3 = e0.elapsed_time(e1) / reps
    out["adam_300"] = {"ms": ms3, "parameters": n3, "GB/s": 28 * n3 / (ms3 * 1e-3) / 1e9, "frac_of_hbm_peak": 28 * n3 / (ms3 * 1e-3) / 1e9 / peak_gbs,
                       "note": "factors of the 300^3 grid: 485 MB per step, larger than L2 (back-to-back launches, no graph)"}
    del call3, p3, g3
    G = params["density_vector"].shape[-1]
    G2 = int(round(G * 1.27))
    moved = 0
    outs, scr = {}, None
    for which in ("density", "appearance"):
        v, m = params[f"{which}_vector"], params[f"{which}_matrix"]
        moved += 4 * (v.numel() + m.numel()) * (1 + (G2 / G) ** 2)
        outs[which] = (torch.empty((3, v.shape[1], G2), device=dev), torch.empty((3, v.shape[1], G2, G2), device=dev))
        nb = ops.vm_resize_scratch_bytes(v.shape[1], G, G2)
        scr = torch.empty(max(nb, scr.numel() if scr is not None else 16), dtype=torch.uint8, device=dev)
    for i in range(2):
        ops.vm_resize(params["density_vector"], params["density_matrix"], G2, out=outs["density"], scratch=scr)
    torch.cuda.synchronize()
    e0.record()
    for i in range(reps):
        for which in ("density", "appearance"):
            ops.vm_resize(params[f"{which}_vector"], params[f"{which}_matrix"], G2, out=outs[which], scratch=scr)
    e1.record()
    torch.cuda.synchronize()
    rs_ms = e0.elapsed_time(e1) / reps
    out["resize"] = {"ms": rs_ms, "from": G, "to": G2, "GB/s": moved / (rs_ms * 1e-3) / 1e9,
                     "note": "both factor sets; outputs and scratch allocated once (TrainState.resize_grid reuses them for params, mu, nu)"}
    return out


class TrainBench:
    """One sharded training step (render_rays forward + MSE + reverse w.r.t. every leaf + the gradient exchange) of a
    workload, set up once: parameters replicated, this rank's ray shard resident in HBM (and mirrored in one pinned host
    buffer for the end-to-end loop), all gradient leaves in one flat buffer."""

    def __init__(self, args, w, world, rank, dev, dist, exchange_arg, packed=False):
        from tensorf_b200 import dist as tdist, ops
        from tensorf_b200.data import HostStage
        self.ops, self.dist, self.w, self.world, self.rank, self.dev = ops, dist, w, world, rank, dev
        self.R_global = w.R * world
        inp = S.make_inputs(w, seed_rays=1 + rank)
        desc = ops.make_desc(R=w.R, N=w.N, K=w.K, G=w.G, cd=w.cd, ca=w.ca, contracted=w.contracted, feat_freqs=w.feat_freqs,
                             view_freqs=w.view_freqs, num_cameras=w.num_cameras, loss_scale=1.0 / (3 * self.R_global), packed_factors=packed)
        self.desc = desc
        self.call = ops.RenderCall(desc, dev)

        def dv(x):
            t = torch.from_numpy(np.ascontiguousarray(x))
            if t.dtype == torch.uint32:
                t = t.view(torch.int32)
            return t.to(dev)

        self.params = {k: dv(v) for k, v in inp["params"].items()}
        if packed:  # parameters (and gradients) live in the kernel-native texel-major layout: no pack / unpack passes per step
            self.params = ops.pack_params(self.params)
        self.host_keys = ["origins", "directions", "camera_indices", "colors", "jitter", "gumbel"]
        # e2e path: the minibatch lives in ONE pinned host buffer mirrored by one device buffer (data.HostStage), so a
        # step's inputs are a single H2D copy; the device views are what both timed loops read
        self.stage = HostStage({k: inp[k] for k in self.host_keys}, dev)
        self.stage.upload()
        self.dins = dict(self.stage.device)
        self.dins["aabb"] = dv(inp["aabb"])
        if w.contracted:
            # host constants of render.py:127-155 (numpy, float64 -> fp32), computed once
            from tensorf_b200.schedule import contracted_schedule
            base, delta = contracted_schedule(w.near, w.far, w.N)
            self.dins["base_ts"], self.dins["deltas"] = dv(base), dv(delta)
        # all gradient leaves live in ONE flat buffer (with the loss in one extra slot) -> one exchange per step; the
        # reverse pass can run in two halves so that the exchange of everything but the density factors overlaps with
        # the density scatter (tensorf_render_rgb_bwd_phase)
        self.fg = tdist.FlatGrads(ops.param_shapes(desc), dev, loss_slot=True)
        self.grads = self.fg.leaves
        self.overlap = world > 1 and not args.no_overlap
        # exchange over peer memory (tensorf_peer_allreduce: P2P loads/stores or NVSwitch multicast through torch
        # symmetric memory) unless --exchange nccl; every rank must agree, else all fall back to NCCL
        self.peer, self.exchange = None, "nccl"
        if world > 1 and exchange_arg != "nccl":
            try:
                shapes = ops.param_shapes(desc)
                self.peer = tdist.PeerAdam(shapes, {k: 0.0 for k in shapes}, dev, multicast=(exchange_arg == "peer-multicast"))
            except Exception as e:  # no symmetric memory on this box: say so, use NCCL
                print(f"[bench] rank {rank}: peer-memory exchange unavailable ({repr(e)[:300]}); using NCCL", file=sys.stderr)
            ok = torch.tensor([0 if self.peer is None else 1], device=dev)
            dist.all_reduce(ok, op=dist.ReduceOp.MIN)
            if int(ok.item()) == 1:
                self.exchange = "peer-multicast" if self.peer.multicast else "peer-p2p"
                self.grads = self.peer.grads
            else:
                self.peer = None
        # exchange transport (P2P vs NVSwitch multicast) and schedule (one exchange after the reverse pass vs the early
        # bucket on a side stream, a few CTAs wide, beside the density scatter) are MEASURED at start-up under --exchange auto:
        # which one wins depends on the world size and the payload (12.8 MB at 128^3, 69.5 MB at 300^3)
        self.tuning = {}
        if self.peer is not None and exchange_arg == "auto":
            self.tuning["transport_us"] = self.peer.autotune_transport()
            self.exchange = "peer-multicast" if self.peer.multicast else "peer-p2p"
        self.peer_overlap = self.peer is not None and exchange_arg == "peer-overlap" and not args.no_overlap
        self.overlap_ctas = 32
        if self.peer is not None:
            self.side, self.ev_early, self.ev_done = torch.cuda.Stream(device=dev), torch.cuda.Event(), torch.cuda.Event()
        self.flush = None if args.no_l2_flush else torch.empty(256 * 1024 * 1024 // 4, dtype=torch.float32, device=dev)
        self.graph, self.graph_loss = None, None
        self.use_graph = not args.no_graph and (world == 1 or self.peer is not None)
        if self.peer is not None and exchange_arg == "auto" and not args.no_overlap:
            # every schedule is timed the way it will run: captured as a CUDA graph and replayed (eager timings rank them
            # differently: launch gaps hide or expose the barrier launches)
            t, graphs = {}, {}
            for name, sync, ov in (("serial", "barrier", False), ("overlap", "barrier", True), ("serial_kernel_sync", "kernel", False)):
                self.peer.sync, self.peer_overlap = sync, ov
                for _ in range(2):
                    self.step_eager()
                self.barrier()
                g, loss = self._capture_graph() if self.use_graph else (None, None)
                graphs[name] = (g, loss)
                run = g.replay if g is not None else self.step_eager
                for _ in range(2):
                    run()
                self.barrier()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                for _ in range(8):
                    run()
                e1.record()
                self.barrier()
                t[name] = self.max_over_ranks(e0.elapsed_time(e1) / 8)
            best = min(t, key=t.get)
            self.peer.sync = "kernel" if best == "serial_kernel_sync" else "barrier"
            self.peer_overlap = best == "overlap"
            self.graph, self.graph_loss = graphs[best]
            if self.graph is None:
                self.use_graph = False
            self.tuning["schedule_ms"] = t
            self.tuning["timed_as"] = "cuda graph replay" if self.graph is not None else "eager launches"

    def _capture_graph(self):
        """The whole step (12 kernel launches + 1 memset per rank, plus the exchange kernels and their cross-rank barriers at
        N > 1) as ONE CUDA graph - the C ABI only enqueues on the caller's stream, so it is capturable as is (stage timers
        off).  Returns (graph, loss tensor), or (None, None) on every rank if any rank cannot capture (e.g. a
        symmetric-memory barrier that refuses stream capture)."""
        ok, g, loss = 1, None, None
        try:
            cap = torch.cuda.Stream(device=self.dev)
            cap.wait_stream(torch.cuda.current_stream())
            g = torch.cuda.CUDAGraph()
            with torch.cuda.stream(cap):
                self.step_eager()  # warm the capture stream (lazy per-function attributes)
                torch.cuda.synchronize()
                self.barrier()
                with torch.cuda.graph(g, stream=cap):
                    loss = self.step_eager()
            torch.cuda.current_stream().wait_stream(cap)
            torch.cuda.synchronize()
        except Exception as e:  # noqa: BLE001
            print(f"[bench] rank {self.rank}: CUDA-graph capture failed ({repr(e)[:300]}); eager launches", file=sys.stderr)
            ok = 0
        if self.world > 1:
            t = torch.tensor([ok], device=self.dev)
            self.dist.all_reduce(t, op=self.dist.ReduceOp.MIN)
            ok = int(t.item())
        return (g, loss) if ok else (None, None)

    def capture(self):
        """Captures the step once (unless the schedule tuning already did); the timed loops replay the graph."""
        if not self.use_graph or self.graph is not None:
            return
        self.graph, self.graph_loss = self._capture_graph()
        if self.graph is None:
            self.use_graph = False

    def describe_exchange(self):
        if self.world == 1:
            return "single GPU"
        if self.peer is not None:
            how = "in two buckets, the early one on a side stream beside the density scatter" if self.peer_overlap else "on the launch stream"
            return f"rays sharded x{self.world}, gradient all-reduce by tensorf_peer_allreduce ({self.exchange}, {self.peer.sync} sync) {how}"
        return f"rays sharded x{self.world}, NCCL grad allreduce" + (" in two buckets overlapped with the density scatter" if self.overlap else "")

    def step(self):
        if self.graph is not None:
            self.graph.replay()
            return self.graph_loss
        return self.step_eager()

    def step_eager(self):
        call, params, dins, grads, peer, fg = self.call, self.params, self.dins, self.grads, self.peer, self.fg
        if self.peer_overlap:
            rgb, loss = call.forward(params, dins, loss_out=peer.loss)
            call.backward(None, grads, phase=1)
            self.ev_early.record()
            with torch.cuda.stream(self.side):
                self.side.wait_event(self.ev_early)
                peer.allreduce("early", channel=1, max_ctas=self.overlap_ctas)
                self.ev_done.record()
            call.backward(None, grads, phase=2)
            peer.allreduce("late", channel=0)
            torch.cuda.current_stream().wait_event(self.ev_done)
            return loss
        if peer is not None:
            rgb, loss = call.forward(params, dins, loss_out=peer.loss)
            call.backward(None, grads)
            peer.allreduce()
            return loss
        rgb, loss = call.forward(params, dins, loss_out=fg.loss)
        if self.overlap:
            call.backward(None, grads, phase=1)
            fg.start_allreduce("early")
            call.backward(None, grads, phase=2)
            fg.start_allreduce("late")
            fg.finish()
        else:
            call.backward(None, grads)
            fg.allreduce()
        return loss

    def barrier(self):
        if self.world > 1:
            self.dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(self, ms):
        t = torch.tensor([ms], dtype=torch.float64, device=self.dev)
        if self.world > 1:
            self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return float(t.item())

    def timed(self, steps, warmup):
        """`warmup` untimed steps, then exactly `steps` steps: CUDA events per step, L2 flushed between steps, barrier +
        synchronize on both sides, max over ranks.  Returns (ms_per_step, launches, {stage: (total ms, calls)}, total ms)."""
        ops = self.ops
        for _ in range(warmup):
            self.step_eager()
        self.barrier()
        prof, launches_per_step = None, None
        if self.use_graph:
            # per-stage times and the launch count come from a short EAGER loop with the library's stage events on (they
            # cannot be captured); the timed region below replays the captured graph
            ops.profile_enable(True)
            l0 = ops.launch_count()
            for _ in range(5):
                if self.flush is not None:
                    self.flush.zero_()
                self.step_eager()
            torch.cuda.synchronize()
            launches_per_step = (ops.launch_count() - l0) // 5
            prof = ops.profile_read()
            ops.profile_enable(False)
            self.capture()
            for _ in range(3):
                self.step()
            self.barrier()
        if not self.use_graph:
            ops.profile_enable(True)
        launches0 = ops.launch_count()
        ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
        self.barrier()
        for i in range(steps):
            if self.flush is not None:
                self.flush.zero_()
            ev[i][0].record()
            self.step()
            ev[i][1].record()
        self.barrier()
        if self.use_graph:
            launches = launches_per_step * steps  # kernels inside the replayed graphs
        else:
            launches = ops.launch_count() - launches0
            prof = ops.profile_read()
            ops.profile_enable(False)
        total_ms = self.max_over_ranks(sum(a.elapsed_time(b) for a, b in ev))
        return total_ms / steps, launches, prof, total_ms

    def timed_e2e(self, steps):
        """The same step through the public API with HOST inputs: one pinned H2D copy of the minibatch and a blocking read
        of the loss every step, inside the timed region."""
        self.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        loss_host = 0.0
        for i in range(steps):
            self.stage.upload()
            loss = self.step()
            loss_host = float(loss.item())  # blocking D2H read, like training.py:342
        e1.record()
        self.barrier()
        h2d = sum(self.stage.host[k].numel() * self.stage.host[k].element_size() for k in self.host_keys)
        return self.max_over_ranks(e0.elapsed_time(e1)) / steps, h2d, loss_host


def multi_rank_parity(ops, dev, world, rank, dist):
    """N > 1, outside every timed region: the sharded path against one rank doing the whole batch.  A small lego-type
    global batch (64 rays per rank) is differentiated two ways - every rank its shard + `tensorf_peer_allreduce`
    (k_peer_allreduce), and rank 0 alone over the concatenated rays - and one optimiser step is taken two ways -
    `tensorf_adam_step_peer` (k_adam_peer: reduce-scatter + Adam on the owner's shard + all-gather) and
    `tensorf_adam_step` on the single-rank gradient.  Reports the largest relative differences over the leaves."""
    from tensorf_b200 import dist as tdist
    Rl = 64
    w = S.Workload("parity", Rl * world, 24, 8, 16, 60, 9, 2, 2)
    inp = S.make_inputs(w, seed_rays=7)  # same global batch on every rank

    def dv(x):
        t = torch.from_numpy(np.ascontiguousarray(x))
        return (t.view(torch.int32) if t.dtype == torch.uint32 else t).to(dev)

    shapes = None
    out = {}
    try:
        desc_l = ops.make_desc(R=Rl, N=w.N, K=w.K, G=w.G, cd=w.cd, ca=w.ca, feat_freqs=2, view_freqs=2, loss_scale=1.0 / (3 * w.R))
        shapes = ops.param_shapes(desc_l)
        lrs = {k: -(0.02 if k.startswith(("density_", "appearance_")) else 1e-3) for k in shapes}
        peer = tdist.PeerAdam(shapes, lrs, dev)
    except Exception as e:
        return {"unavailable": repr(e)[:200]}
    params = {k: dv(v) for k, v in inp["params"].items()}
    peer.load_params(params)
    a, b = rank * Rl, (rank + 1) * Rl
    per_ray = ("origins", "directions", "camera_indices", "colors")
    dins = {k: dv(inp[k][a:b] if k in per_ray else inp[k]) for k in per_ray + ("jitter", "gumbel", "aabb")}
    call = ops.RenderCall(desc_l, dev)
    call.forward(peer.params, dins, loss_out=peer.loss)
    call.backward(None, peer.grads)
    local = {k: v.clone() for k, v in peer.grads.items()}
    peer.allreduce()
    torch.cuda.synchronize()
    summed = {k: v.clone() for k, v in peer.grads.items()}
    # fused exchange + Adam from the LOCAL gradients again
    for k in local:
        peer.grads[k].copy_(local[k])
    peer.step(count=0)
    torch.cuda.synchronize()
    new_params = {k: v.clone() for k, v in peer.params.items()}
    if rank == 0:
        desc_g = ops.make_desc(R=w.R, N=w.N, K=w.K, G=w.G, cd=w.cd, ca=w.ca, feat_freqs=2, view_freqs=2, loss_scale=1.0 / (3 * w.R))
        full = ops.RenderCall(desc_g, dev)
        fins = {k: dv(inp[k]) for k in per_ray + ("jitter", "gumbel", "aabb")}
        full.forward(params, fins, loss_out=torch.zeros(1, device=dev))
        g1 = full.backward(None)
        rel_inf = rel_l2 = 0.0
        for k in g1:
            ref, x = g1[k].double(), summed[k].double()
            rel_inf = max(rel_inf, float((x - ref).abs().max() / ref.abs().max().clamp_min(1e-300)))
            rel_l2 = max(rel_l2, float((x - ref).norm() / ref.norm().clamp_min(1e-300)))
        names = list(shapes)
        p1 = [params[k].clone() for k in names]
        adam = ops.AdamCall(p1, [torch.zeros_like(x) for x in p1], [torch.zeros_like(x) for x in p1], [lrs[k] for k in names])
        adam.step([g1[k] for k in names], count=0)
        torch.cuda.synchronize()
        step_inf = step_l2 = 0.0
        for k, x in zip(names, p1):
            upd_ref, upd = (x - params[k]).double(), (new_params[k] - params[k]).double()
            step_inf = max(step_inf, float((upd - upd_ref).abs().max() / upd_ref.abs().max().clamp_min(1e-300)))
            step_l2 = max(step_l2, float((upd - upd_ref).norm() / upd_ref.norm().clamp_min(1e-300)))
        out = {"max_rel_inf": rel_inf, "max_rel_l2": rel_l2, "adam_update_max_rel_inf": step_inf, "adam_update_max_rel_l2": step_l2, "rays_global": w.R, "ranks": world,
               "covers": "k_peer_allreduce (summed gradient) and k_adam_peer (parameter update) vs one rank over the concatenated rays",
               "note": "sums are re-associated across ranks: gradient differences are fp32 rounding of the ray partition; the first Adam "
                       "update is lr * g / (|g| + 1e-8), which amplifies that rounding for the few elements with |g| ~ 1e-8 (inf norm) "
                       "and not otherwise (L2 norm)"}
    del peer
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="lego_256")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-l2-flush", action="store_true")
    ap.add_argument("--no-render", action="store_true")
    ap.add_argument("--no-config3", action="store_true", help="skip the BASELINE configs[2] sub-record (300^3, 16384 global rays, strong scaling)")
    ap.add_argument("--no-packed", action="store_true", help="skip the packed-factor-layout sub-record")
    ap.add_argument("--no-graph", action="store_true", help="N=1: launch every kernel of a step from the host instead of replaying one CUDA graph")
    ap.add_argument("--no-overlap", action="store_true", help="N>1: one exchange after the whole reverse pass")
    ap.add_argument("--exchange", default="auto", choices=["auto", "peer-p2p", "peer-multicast", "peer-overlap", "nccl"],
                    help="N>1 gradient exchange: own kernel over peer memory (auto = peer-overlap: two buckets, the early one beside the "
                         "density scatter; NCCL if symmetric memory is unavailable), peer-p2p / peer-multicast = one exchange on the launch stream, or NCCL")
    args = ap.parse_args()
    w = workload_from_name(args.workload)
    if args.impl == "reference":
        return run_reference_arm(args, w)

    import torch.distributed as dist
    from tensorf_b200 import ops

    # rank 0 prints exactly ONE JSON line on stdout: anything libraries write to fd 1 (e.g. NCCL's
    # version banner) is sent to stderr instead
    sys.stdout.flush()
    json_fd = os.dup(1)
    os.dup2(2, 1)

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus and world > 1:
        raise SystemExit(f"--gpus {args.gpus} but WORLD_SIZE={world}")
    assert torch.cuda.is_available(), "bench.py needs a GPU (there is no CPU fallback)"
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    warmup = max(args.warmup, 3)
    steps = args.steps
    dist_or_none = dist if world > 1 else None

    tb = TrainBench(args, w, world, rank, dev, dist, args.exchange)
    R_global = tb.R_global
    # clocks are sampled from the warm-up through the timed and end-to-end loops (the timed region
    # alone is only tens of milliseconds, shorter than nvidia-smi's sampling period)
    sampler = ClockSampler(local) if rank == 0 else None
    ms_per_step, launches, prof, total_ms = tb.timed(steps, warmup)
    value = R_global / (ms_per_step * 1e-3)
    e2e_ms, h2d, loss_host = tb.timed_e2e(steps)
    e2e_value = R_global / (e2e_ms * 1e-3)
    clocks = sampler.stop() if sampler else None
    exchange_desc = tb.describe_exchange()
    params, grads = tb.params, tb.grads

    # ---- BASELINE configs[2]: upsampled 300^3 grid, 16384-ray global batch strong-scaled over the ranks ----------------
    config3 = None
    if not args.no_config3 and args.workload == "lego_256":
        Rg = 16384
        w3 = S.lego_workload(R=Rg // world, G=300, name=f"lego_G300_Rglobal{Rg}_N519_K77 (BASELINE configs[2])")
        tb3 = TrainBench(args, w3, world, rank, dev, dist, args.exchange)
        ms3, _, prof3, tot3 = tb3.timed(max(4, steps // 2), 3)
        st3 = {k: round(v[0] / max(v[1], 1), 4) for k, v in sorted(prof3.items(), key=lambda kv: -kv[1][0])}
        n_par = sum(int(np.prod(s_)) for s_ in ops.param_shapes(tb3.desc).values())
        config3 = {"workload": w3.name, "scaling": "strong", "R_global": Rg, "R_per_gpu": w3.R, "N": w3.N, "K": w3.K, "G": 300,
                   "value": Rg / (ms3 * 1e-3), "unit": UNIT, "ms_per_step": ms3, "exchange_bytes": 4 * n_par, "exchange": tb3.describe_exchange(),
                   "exchange_tuning": tb3.tuning, "cuda_graph": tb3.graph is not None,
                   "stages_ms": st3,
                   "roofline_step_frac": Rg / world / (ms3 * 1e-3) * w3.train_bytes_per_ray() / 1e9 / measured_peaks()[0]}
        del tb3
        torch.cuda.empty_cache()

    # ---- the same step with parameters / gradients kept in the kernel-native packed layout (TENSORF_FLAG_PACKED_FACTORS) ----
    packed_rec = None
    if not args.no_packed:
        tbp = TrainBench(args, w, world, rank, dev, dist, args.exchange, packed=True)
        msp, _, profp, _ = tbp.timed(max(6, steps // 2), 3)
        packed_rec = {"value": R_global / (msp * 1e-3), "unit": UNIT, "ms_per_step": msp,
                      "stages_ms": {k: round(v[0] / max(v[1], 1), 4) for k, v in sorted(profp.items(), key=lambda kv: -kv[1][0])},
                      "note": "TENSORF_FLAG_PACKED_FACTORS: factors, their gradients (and, in a training loop, the Adam moments - every optimiser "
                              "operation is elementwise) stay in the texel-major layout the kernels read; the per-step pack and unpack passes "
                              "disappear.  The headline `value` keeps the reference's channel-first layout at the boundary."}
        del tbp
        torch.cuda.empty_cache()

    parity = multi_rank_parity(ops, dev, world, rank, dist) if world > 1 else None

    # ---- render_360 as a job (BASELINE configs[4] shapes), this rank's stripes of every frame ----
    render = None
    if not args.no_render:
        del tb.call
        torch.cuda.empty_cache()
        render = bench_render(ops, dev, world, rank, dist_or_none)

    if rank == 0:
        peak, tc_peak, peak_src = measured_peaks()
        # Algorithmic bytes of the gather model: 6 taps x 4 B per sample-channel,
        # the reverse pass re-gathers and scatter-adds the same count (2x).
        alg = {
            "density_select": 24 * w.N * 3 * w.cd * w.R,
            "appearance_gather": 24 * w.K * 3 * w.ca * w.R,
            "density_scatter": 48 * w.N * 3 * w.cd * w.R,
            "appearance_scatter": 48 * w.K * 3 * w.ca * w.R,
        }
        # Algorithmic flops of the MLP (networks.py:57-117): 2 x (3ca*27 + enc*128 + 128*128 + 128*3) per row forward,
        # twice that in the reverse pass (activation and weight gradients).
        enc = w.encoded_dim()
        mlp_row = 2 * (3 * w.ca * 27 + enc * 128 + 128 * 128 + 128 * 3)
        flops = {"mlp_fwd": mlp_row * w.R * w.K, "mlp_bwd": 2 * mlp_row * w.R * w.K}
        stages = {k: {"ms": v[0] / max(v[1], 1), "share": v[0] / max(v[1], 1) / ms_per_step} for k, v in prof.items()}
        traffic_all = {}
        tp = ROOT / "profiles" / "traffic.json"  # per-step ncu numbers of the committed --set full capture (tools/traffic_from_ncu.py)
        if tp.exists():
            traffic_all = json.loads(tp.read_text()).get(args.workload, {})
        # per-stage rooflines: gather / scatter stages against HBM (gather-model bytes; DRAM and L2 rates from the ncu
        # capture's bytes over the LIVE stage time), MLP stages against the tensor pipe
        stage_roofs = {}
        for k, st_ in stages.items():
            tr = traffic_all.get(k) if isinstance(traffic_all.get(k), dict) else None
            e = {"ms": round(st_["ms"], 4)}
            if k in alg:
                e.update(bound="hbm", algorithmic_gbs=alg[k] / (st_["ms"] * 1e-3) / 1e9)
            if k in flops:
                e.update(bound="tensor", algorithmic_tflops=flops[k] / (st_["ms"] * 1e-3) / 1e12,
                         frac_of_bf16_sustained=flops[k] / (st_["ms"] * 1e-3) / 1e12 / tc_peak)
            if tr:
                e.update(dram_gbs=tr["dram_bytes"] / (st_["ms"] * 1e-3) / 1e9, lts_gbs=tr["l2_bytes"] / (st_["ms"] * 1e-3) / 1e9,
                         dram_frac_of_hbm_peak=tr["dram_bytes"] / (st_["ms"] * 1e-3) / 1e9 / peak)
            stage_roofs[k] = e
        dom = max(stages, key=lambda k: stages[k]["ms"])  # the dominant stage over ALL stages
        dom_ms = stages[dom]["ms"]
        dtr = traffic_all.get(dom) if isinstance(traffic_all.get(dom), dict) else None
        traffic = dtr["dram_bytes"] if dtr else None
        if dom in flops:
            achieved = flops[dom] / (dom_ms * 1e-3) / 1e12
            roofline = {"bound": "tensor", "kernel": dom, "achieved": achieved, "peak": tc_peak, "unit": "TFLOP/s", "frac": achieved / tc_peak,
                        "traffic": traffic, "algorithmic_flops_per_launch": flops[dom], "avg_launch_ms": dom_ms, "peak_source": peak_src,
                        "split_factor": 3,
                        "note": "algorithmic fp32 flops of the FeatureMlp stage (all its kernels) over the measured dense bf16 peak; fp32 parity on "
                                "16-bit tensor cores costs 3 tcgen05.mma per product (two-term operand split), so the issued-MMA fraction is 3x frac"}
        else:
            achieved = alg.get(dom, 0) / (dom_ms * 1e-3) / 1e9
            roofline = {"bound": "hbm", "kernel": dom, "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                        "traffic": traffic, "algorithmic_bytes_per_launch": alg.get(dom), "avg_launch_ms": dom_ms, "peak_source": peak_src,
                        "note": "gather-model bytes (6 taps x 4 B per sample-channel, SURVEY 8d); at 128^3 the factors and their gradients are "
                                "L2-resident, so frac > 1 is an L2/LSU rate - `traffic` is the stage's ncu DRAM bytes per step"}
        out = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": steps, "warmup": warmup,
            "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": {"workload": w.name, "R_per_gpu": w.R, "R_global": R_global, "N": w.N, "K": w.K, "G": w.G,
                       "cd": w.cd, "ca": w.ca, "feat_freqs": w.feat_freqs, "view_freqs": w.view_freqs,
                       "contracted": w.contracted, "parallelism": exchange_desc, "exchange_tuning": tb.tuning,
                       "l2": "flushed (256 MiB write) between timed steps" if tb.flush is not None else "not flushed",
                       "timed": "render_rays fwd + MSE + reverse wrt all LearnableParams leaves; Adam excluded" +
                                ("; the step is captured once and replayed as ONE CUDA graph per rank (stages_ms from a separate eager loop)" if tb.graph is not None else ""),
                       "launch_gap_frac": 1.0 - sum(v_[0] / max(v_[1], 1) for v_ in prof.values()) / ms_per_step,
                       "loss": loss_host},
            "roofline_step": {"bound": "hbm", "achieved": value / world * w.train_bytes_per_ray() / 1e9, "peak": peak, "unit": "GB/s",
                              "frac": value / world * w.train_bytes_per_ray() / 1e9 / peak,
                              "note": "whole step per GPU, gather-model bytes 3*24*(N*3cd+K*3ca) per ray (SURVEY §8d); factors are L2-resident at 128^3"},
            "roofline": roofline,
            "stage_rooflines": stage_roofs,
            "stages_ms": {k: round(v["ms"], 4) for k, v in sorted(stages.items(), key=lambda kv: -kv[1]["ms"])},
            "e2e": {"value": e2e_value, "unit": UNIT, "ms_per_step": e2e_ms, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": 4,
                    "note": "public API (RenderCall + data.HostStage): one pinned H2D copy of the minibatch and a blocking loss read per "
                            "step; no L2 flush in this loop (the inputs arrive from the host every step), hence ~= the flushed device-timed value"},
            "gpu_launches": int(launches),
            "clocks": clocks,
        }
        if packed_rec is not None:
            out["packed_layout"] = packed_rec
        if config3 is not None:
            out["config3"] = config3
        if parity is not None:
            out["parity_check"] = parity
        if render is not None:
            out["render"] = render
        if not args.no_render:
            out["optimizer"] = bench_optimizer(ops, dev, params, grads, peak)
        if world == 1 and not args.no_cpu_baseline:
            v, ms, cores = cpu_reference_rays_per_s(w, 512, 6, 1)
            out["cpu_baseline"] = {"value": v, "unit": UNIT, "cores": cores, "kind": "port",
                                   "sample": f"512-ray slice of the same batch, fwd+bwd, 6 timed steps ({ms:.0f} ms each), "
                                             f"PyTorch-CPU fp32 restatement of the reference (JAX not in image)"}
        os.write(json_fd, (json.dumps(out) + "\n").encode())
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
