#!/usr/bin/env python
"""bench.py — headline benchmark of the B200-native NeRF hot path.

    python bench.py --gpus N --steps K --warmup W            # this repo (libb2nerf.so); torchrun launches it for N > 1
    python bench.py --impl reference --gpus N --steps K ...   # the repo's ops as pure PyTorch on the host cores (CPU arm)

Workload at every N (BASELINE.json configs[1] / configs[3]): one step = one BATCH of FRAMES_PER_STEP 512x512 talking-head frames per GPU (three of the
reference's inference-loader batches of 32, provider_for_inference.py:727) — per frame: AudioNet+AudioAttNet, near/far, the 16-step occupancy march /
tri-plane encode / head MLPs / composite loop, background blend — random-init weights, the synthetic head-sized density blob of SURVEY §8d,
`max_steps 16, dt_gamma 1/256, bound 1`.  Frames are independent, so N GPUs render N batches per step (weak scaling, no data-path collective).  A batch per
step keeps the timed region of the driver's 20-step run above half a second at every N (a 20-frame region is 7 ms: its max-over-ranks is launch jitter).

One JSON line on stdout (rank 0).  `value` = frames/s with the rays already in HBM; `e2e` = the same through FramePipeline.submit_host_pose with pinned HOST
buffers (4x4 pose + audio window up, RGB24 frame down, every frame); `roofline` = the fused tcgen05 head kernel (isolated timing -> burst peak), `kernels` =
per-kernel rooflines of the per-op kernels; `cpu_baseline` = oracle/torch_port.py on a bounded, evenly strided sample of the same frame;
`reference_on_b200` = the unmodified reference model on its own CUDA extensions on this GPU (profiles/reference_on_b200.py), when staged.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
for p in (os.path.join(ROOT, "lzzx-nerf_b200"), ROOT):
    if p not in sys.path:
        sys.path.insert(0, p)

import numpy as np  # noqa: E402
import torch  # noqa: E402

HW = 512
N_RAYS = HW * HW
POOL = 24                      # distinct frames cycled through: 24 x 6.3 MB of rays = 151 MB > 126 MB L2
FRAMES_PER_STEP = 96           # one step = 3 inference-loader batches of 32 frames (provider_for_inference.py:727)
ROW_STRIDE = 8                 # CPU arms: every 8th row of the frame (64 of 512 rows, evenly spread: the sample has the whole frame's samples-per-ray mix)
METRIC = "infer_512x512_frames_per_sec"
MACS_PER_SAMPLE_INFER = 23184  # SURVEY §8a a7 (no unc_net at inference)


def peaks():
    try:
        return json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))), "measured"
    except Exception:
        return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0}, "fallback"


def ncu_traffic():
    """dram__bytes_read.sum + dram__bytes_write.sum per launch of the dominant kernel, from the committed ncu --set full capture (profiles/)."""
    try:
        return json.load(open(os.path.join(ROOT, "profiles", "head_kernel_ncu.json")))["dram_bytes_per_launch"]
    except Exception:
        return None


class ClockSampler(threading.Thread):
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""

    def __init__(self, index):
        super().__init__(daemon=True)
        q = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
            "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"
        self.cmd = ["nvidia-smi", f"--id={index}", f"--query-gpu={q}", "--format=csv,noheader,nounits", "-lms", "100"]
        self.rows, self.proc = [], None
        self.cmd[-1] = "50"

    def run(self):
        try:
            self.proc = subprocess.Popen(self.cmd, stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            for line in self.proc.stdout:
                self.rows.append([c.strip() for c in line.split(",")])
        except Exception:
            pass

    def wait_first(self, timeout=10.0):
        t0 = time.time()
        while not self.rows and time.time() - t0 < timeout:
            time.sleep(0.05)

    def stop(self):
        if self.proc is not None:
            self.proc.terminate()
        sm, mx, reasons = [], 0.0, set()
        for r in self.rows:
            try:
                sm.append(float(r[0])); mx = max(mx, float(r[1]))
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[3:7]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
            except Exception:
                continue
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": mx or None, "reasons": sorted(reasons), "samples": len(sm)}


def synthetic_inputs(rank):
    from b2nerf import scene
    frames = [scene.frame_rays(frame=rank * 1000 + f) for f in range(POOL)]
    auds = [scene.audio_window(frame=rank * 1000 + f) for f in range(POOL)]
    bitfield = scene.bitfield_from_grid(scene.density_grid())
    return frames, auds, bitfield


def build_model(device=None):
    from b2nerf.model import HeadModel
    torch.manual_seed(0)
    m = HeadModel(audio_in_dim=1024)       # random-init weights of the reference architecture (683 509 parameters)
    m.testing = True
    return m if device is None else m.to(device)


# ------------------------------------------------------------------------------------------------------------------------------
# CPU arm: the repo's ops as pure PyTorch on the host cores
# ------------------------------------------------------------------------------------------------------------------------------
def cpu_frames_per_sec(rows, steps, warmup):
    """Times oracle/torch_port.render_frame on `rows` rows of the 512x512 frame taken at an even stride over the whole frame (so the sample has the frame's own mix
    of empty and dense rays); returns (frames/s equivalent, seconds/step, samples)."""
    from oracle import torch_port as tp
    from b2nerf import scene
    torch.set_num_threads(os.cpu_count())
    m = build_model()
    p = tp.params_from_state_dict(m.state_dict())
    bf = torch.from_numpy(scene.bitfield_from_grid(scene.density_grid()))
    sel = torch.arange(rows) * (HW // rows) + (HW // rows) // 2
    aabb = torch.from_numpy(scene.AABB)
    times, ns = [], 0
    with torch.no_grad():
        for s in range(warmup + steps):
            o, d = scene.frame_rays(frame=s)
            o = torch.from_numpy(o).view(HW, HW, 3)[sel].reshape(-1, 3)
            d = torch.from_numpy(d).view(HW, HW, 3)[sel].reshape(-1, 3)
            auds = torch.from_numpy(scene.audio_window(frame=s))
            t0 = time.perf_counter()
            enc_a = m.encode_audio(auds)[0]
            _, _, _, ns = tp.render_frame(p, o, d, bf, enc_a, m.individual_codes[0].detach(), torch.tensor([0.4]), aabb)
            dt = time.perf_counter() - t0
            if s >= warmup:
                times.append(dt)
    sec = float(np.mean(times))
    return (rows / HW) / sec, sec, ns


def run_reference_arm(args, rank):
    if rank != 0:
        return
    # bounded sample per step: every 8th row of the frame (64 rows), thinned further (powers of two) when --steps is large so that the whole run stays within ~3 minutes
    rows = HW // ROW_STRIDE
    _, probe_sec, _ = cpu_frames_per_sec(rows, 1, 1)
    budget = 180.0 / max(1, args.steps + max(args.warmup, 1))
    while probe_sec * rows / (HW // ROW_STRIDE) > budget and rows > 8:
        rows //= 2
    fps, sec, ns = cpu_frames_per_sec(rows, args.steps, max(args.warmup, 1))
    cores = os.cpu_count()
    sample = (f"{rows} rows of the 512x512 frame at an even stride of {HW // rows} ({rows * HW} of {N_RAYS} rays, {ns} samples, x{HW // rows} extrapolation) per step, "
              f"pure-PyTorch fp32 port (oracle/torch_port.py)")
    line = {"impl": "reference", "metric": METRIC, "value": fps, "unit": "frames/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": sec * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": "infer_512x512_frame", "sample_rows": rows, "extrapolation_factor": HW // rows,
                       "note": "the reference has no CPU path for these ops (SURVEY §8d); this arm is the pure-PyTorch port on host cores, timed on an evenly strided "
                               "subset of the frame's rows and scaled by rows"},
            "cpu_baseline": {"value": fps, "unit": "frames/s", "cores": cores, "kind": "port", "sample": sample},
            "e2e": {"value": fps, "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    _emit(line)


# ------------------------------------------------------------------------------------------------------------------------------
# GPU arm
# ------------------------------------------------------------------------------------------------------------------------------
def head_kernel_profile(model, renderer, frames, auds, n_frames, repeats=5):
    """Separate instrumented pass (no CUDA graph; CUDA events cannot be recorded inside a replayed graph).  For each of `n_frames` frames the
    reference loop is run once on the per-op kernels; every loop iteration's fused-head launch is then timed in isolation: one CUDA-event pair on the
    launching stream around `repeats` back-to-back launches of that same call, replayed from a CUDA graph (so no host gap is inside the pair).  Returns (sum of average launch
    durations in ms, launches, samples evaluated)."""
    import raymarching
    stream = torch.cuda.current_stream()
    dev, N, kw = renderer.dev, renderer.N, renderer.kw
    total_ms, launches, samples = 0.0, 0, 0
    with torch.no_grad():
        for s in range(n_frames):
            rays_o, rays_d = frames[s % POOL]
            with torch.autocast("cuda", dtype=torch.float16):
                enc_a = model.encode_audio(auds[s % POOL]).float()
            nears, fars = raymarching.near_far_from_aabb(rays_o, rays_d, model.aabb_infer, 0.05)
            ws, depth, image = torch.zeros(N, device=dev), torch.zeros(N, device=dev), torch.zeros(N, 3, device=dev)
            sa, se, su = torch.zeros(N, device=dev), torch.zeros(N, device=dev), torch.zeros(N, device=dev)
            alive = torch.arange(N, dtype=torch.int32, device=dev); rays_t = nears.clone()
            step = 0
            while step < kw["max_steps"]:
                n_alive = alive.shape[0]
                if n_alive <= 0:
                    break
                n_step = max(min(N // n_alive, 8), 1)
                xyzs, dirs, deltas = raymarching.march_rays(n_alive, n_step, alive, rays_t, rays_o, rays_d, model.bound, model.density_bitfield, model.cascade,
                                                            model.grid_size, nears, fars, 128, False, kw["dt_gamma"], kw["max_steps"])
                sig, rgb, aa, ae, un = model(xyzs, dirs, enc_a, renderer.ind_code, renderer.eye)
                m_eval = n_alive * n_step                   # what b2n_render_frame evaluates (n_valid = n_alive * n_step)
                nv = torch.tensor([m_eval], dtype=torch.int32, device=dev)
                torch.cuda.synchronize(dev)
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                # the repeats are replayed from a small CUDA graph: launched one by one from Python, the ~30 us of host time per call would sit inside the
                # event pair whenever the kernel is shorter than that (the late, small loop iterations) and the host is busy
                g = torch.cuda.CUDAGraph()
                with torch.cuda.graph(g):
                    for _ in range(repeats):
                        model(xyzs, dirs, enc_a, renderer.ind_code, renderer.eye, n_valid=nv, out=(sig, rgb, aa, ae, un))
                stream = torch.cuda.current_stream()
                g.replay()                                  # warm replay (graph upload)
                torch.cuda.synchronize(dev)
                e0.record(stream)
                g.replay()
                e1.record(stream)
                e1.synchronize()
                del g
                total_ms += e0.elapsed_time(e1) / repeats; launches += 1; samples += m_eval
                raymarching.composite_rays_triplane(n_alive, n_step, alive, rays_t, sig, rgb, deltas, aa, ae, un, ws, depth, image, sa, se, su, kw["T_thresh"])
                alive = alive[alive >= 0]
                step += n_step
    return total_ms, launches, samples


def train_kernel_breakdown(step_fn, batches, m_buf, n_rays):
    """Per-kernel durations of ONE replayed training step (CUPTI kernel records via torch.profiler — CUDA events cannot be recorded inside a replayed graph) with
    the algorithmic work of each of the big kernels, so the step's time can be read against HBM / tensor peaks.  Explanatory: the step time itself is the
    event-timed figure above."""
    from torch.profiler import profile, ProfilerActivity
    pk, _ = peaks()
    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        for s in range(2):
            b = batches[s % 4]
            step_fn(b[0], b[1], b[2], b[3], index=1 + s, face_mask=b[4])
        torch.cuda.synchronize()
    ev = sorted((e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA), key=lambda e: e.time_range.start)
    starts = [i for i, e in enumerate(ev) if "k_near_far" in e.name]
    ev = ev[starts[-1]:] if starts else ev
    t0, t1 = ev[0].time_range.start, max(e.time_range.end for e in ev)
    dur = {}
    for e in ev:
        name = e.name.split("(")[0].replace("void ", "").replace("b2n::", "")
        dur[name] = dur.get(name, 0.0) + (e.time_range.end - e.time_range.start)
    M = m_buf
    # algorithmic work per sample of the 327 k-sample buffer: bytes the kernel has to move at least, FLOPs of the layers it evaluates
    work = {
        "k_head_forward<true, true>": dict(flop=2 * 24368 * M, bytes=(24 + 1088 + 28) * M, note="coordinates in, 1088 B of kept fp16 activations + outputs out per sample"),
        "k_head_backward": dict(flop=2 * 2 * 24368 * M, bytes=(744 + 1000 + 144) * M, note="kept activations in, fp16 layer gradients + fp32 table-gradient planes out"),
        "k_linear_wgrad_multi": dict(flop=2 * 24368 * M, bytes=2132 * M, note="every (dY, X) operand pair once"),
        "k_triplane_bwd_fix": dict(bytes=(144 + 12) * M, reds=144 * M, note="144 table reductions per sample"),
        "k_march_train_count<1>": dict(bytes=40 * n_rays, note="instruction-bound DDA (DESIGN §4.5)"),
        "k_march_train_emit": dict(bytes=12 * n_rays + 32 * M),
        "k_comp_train_fwd<1, 2, true>": dict(bytes=36 * M + 44 * n_rays),
        "k_comp_train_bwd<1, 2, true>": dict(bytes=64 * M + 72 * n_rays),
    }
    rows = []
    for name, us in sorted(dur.items(), key=lambda kv: -kv[1]):
        row = {"kernel": name, "us": round(us, 1)}
        w = work.get(name)
        if w:
            row["frac_of_hbm"] = round(w["bytes"] / (us * 1e-6) / 1e9 / pk["hbm_gbs"], 3)
            if "flop" in w:
                row["frac_of_bf16_burst"] = round(w["flop"] / (us * 1e-6) / 1e12 / pk["bf16_tflops"], 3)
            if "note" in w:
                row["note"] = w["note"]
        rows.append(row)
        if len(rows) >= 14:
            break
    return {"how": "CUPTI kernel records of one replayed step (torch.profiler); branches of the graph overlap, so the rows add up to more than the span",
            "span_us": round(t1 - t0, 1), "kernel_sum_us": round(sum(dur.values()), 1), "launches": len(ev), "rows": rows}


def train_bench(dev, rank, world, bitfield, barrier, steps, warmup, n_rays=65536, eager=False, fused_head=True, kernel_breakdown=False):
    """BASELINE configs[2]/[4]: data-parallel training step, 65 536 rays per GPU, synthetic audio window (AudioNet + AudioAttNet), grid backward,
    AdamW; gradients all-reduced once per step over the flat buffer when world > 1.  Returns a dict (rays/s over all ranks)."""
    from b2nerf import scene
    from b2nerf.train import Trainer
    model = build_model(dev)
    model.testing = False
    model.density_bitfield.copy_(torch.from_numpy(bitfield).to(dev))
    tr = Trainer(model, fp16=True, fused_head=fused_head)
    batches = []
    for s in range(4):
        o, d = scene.train_rays(step=rank * 100 + s, n=n_rays)
        batches.append((torch.from_numpy(o).to(dev), torch.from_numpy(d).to(dev), torch.from_numpy(scene.audio_window(rank * 100 + s)).to(dev),
                        torch.rand(n_rays, 3, device=dev), torch.rand(n_rays, device=dev) < 0.5))         # rays, audio window, target colours, face mask
    for s in range(warmup):
        b = batches[s % 4]
        tr.train_step(b[0], b[1], b[2], b[3], index=s, face_mask=b[4])
        if s == 15:
            tr.update_mean_count()            # the reference's warm-up: 16 steps with worst-case buffers, then the mean_count estimate
    step_fn = tr.train_step if eager else tr.train_step_graphed
    for s in range(34):                   # graph captures (the plain step and the every-16th-step variant with the smoothness regulariser) + steady state
        b = batches[s % 4]
        step_fn(b[0], b[1], b[2], b[3], index=s, face_mask=b[4])
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for s in range(steps):
        b = batches[s % 4]
        loss, m_buf = step_fn(b[0], b[1], b[2], b[3], index=s, face_mask=b[4])
    e1.record()
    barrier()
    ms = e0.elapsed_time(e1)
    if world > 1:
        import torch.distributed as dist
        t = torch.tensor([ms], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    kernels = None
    if kernel_breakdown and not eager and rank == 0:
        try:
            kernels = train_kernel_breakdown(step_fn, batches, int(m_buf), n_rays)
        except Exception as e:      # explanatory leg only: never fail the line for it
            kernels = {"error": f"{type(e).__name__}: {e}"}
    return {"metric": "train_rays_per_sec", "value": world * steps * n_rays / (ms * 1e-3), "unit": "rays/s", "ms_per_step": ms / steps, "rays_per_gpu": n_rays,
            "kernels": kernels, "samples_per_step_buffer": int(m_buf), "loss": float(loss), "grad_allreduce_bytes": tr.grads.nbytes() if world > 1 else 0, "steps_timed": steps,
            "grad_allreduce": "none (1 GPU)" if world == 1 else ("one two-shot kernel per rank over NVLink peer memory (csrc/peer_allreduce.cu), inside the step's graph"
                                                              if tr.grads.peer is not None else "ncclAllReduce of the flat buffer inside the step's graph"),
            "objective": "TrainerUtil.py:238-363 head branch: uncertainty-weighted MSE + loss_u + static-uncertainty + entropy(1e-4) + masked / ramped ambient terms, "
                         "smoothness regulariser on every 16th step (two more network passes); AdamW groups of network.py:332-356, LambdaLR stepped every iteration",
            "path": ("eager: " if eager else "forward + backward + all-reduce replayed from one CUDA graph: ") +
                    ("march / composite ops + fused head forward (activations kept) + fused head backward-data + tcgen05 weight-gradient kernel + privatised grid backward"
                     if fused_head else "drop-in ops + torch autograd (MLPs via cuBLAS under autocast fp16)") + ", flat AdamW, flat-buffer NCCL all-reduce"}


def torso_bench(dev, steps=20, warmup=5, head=None):
    """Torso branch of a 512x512 frame (torso is OFF in the headline frame, BASELINE configs[1]): the fused kernel (csrc/fused_torso.cu)
    next to the reference's op-by-op graph on the drop-in encoders + torch Linear under autocast, same synthetic torso occupancy (a blob over ~1/3 of
    the image), CUDA events on the launching stream."""
    import math
    from b2nerf.torso import TorsoModel, get_bg_coords
    torch.manual_seed(0)
    m = TorsoModel().to(dev).eval()
    m.torso_encoder.embeddings.data.uniform_(-0.5, 0.5)
    G = m.grid_size
    yy, xx = torch.meshgrid(torch.linspace(-1, 1, G), torch.linspace(-1, 1, G), indexing="ij")
    m.density_grid_torso.copy_((0.05 * torch.exp(-((xx * 1.2) ** 2 + ((yy - 0.5) * 1.5) ** 2) / 0.3)).reshape(-1).to(dev))
    m.mean_density_torso = float(m.density_grid_torso.mean())
    coords = get_bg_coords(HW, HW, dev)
    pose = torch.eye(4, device=dev)[None].clone(); pose[0, 2, 3] = 3.35
    hc = m.frame_constants(pose, 0)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)

    def timed(fn):
        for _ in range(warmup):
            fn()
        torch.cuda.synchronize(); e0.record()
        for _ in range(steps):
            fn()
        e1.record(); torch.cuda.synchronize()
        return e0.elapsed_time(e1) / steps

    # the fused call is replayed from a CUDA graph (pack + frame kernel): eager, the ~100 us of Python / ctypes per call would be what is timed
    m.run_torso_fused(coords, pose, 0, None, h_const=hc)
    torch.cuda.synchronize()
    graph, side = torch.cuda.CUDAGraph(), torch.cuda.Stream()
    with torch.cuda.stream(side):
        with torch.cuda.graph(graph, stream=side):
            keep = m.run_torso_fused(coords, pose, 0, None, h_const=hc)
    torch.cuda.current_stream().wait_stream(side)
    fused_ms = timed(graph.replay)
    fused_eager_ms = timed(lambda: m.run_torso_fused(coords, pose, 0, None, h_const=hc))
    assert bool(torch.isfinite(keep["bg_color"]).all())
    with torch.no_grad(), torch.autocast("cuda", dtype=torch.float16):
        per_op_ms = timed(lambda: m.run_torso(coords, pose, 0, None))
        n_on = int(m.run_torso(coords, pose, 0, None)["mask"].sum())
    # torso TRAINING stage (opt.torso: MSE on torso_color, TrainerUtil.py:188-236): forward + backward of one 512x512 frame, fused kernels vs autograd over the op graph
    target = torch.rand(HW * HW, 3, device=dev)
    params = [p for p in m.parameters()]

    def train_step(fused):
        for p in params:
            p.grad = None
        with torch.autocast("cuda", dtype=torch.float16):
            res = m.run_torso_train_fused(coords, pose, 5, None) if fused else m.run_torso(coords, pose, 5, None)
            loss = ((res["torso_color"] - target) ** 2).mean()
        (loss * 1024.0).backward()

    m.train()
    train_fused_ms = timed(lambda: train_step(True))
    train_per_op_ms = timed(lambda: train_step(False))
    m.eval()
    macs = 5440                                            # 34*32 + 32*32 + 32*2 + 66*32 + 32*32 + 32*4 per torso pixel (constant inputs folded into a bias)
    frames_with_torso = None
    if head is not None:                                   # whole frames with the torso behind the head: FrameRenderer(torso=...) on every pipeline slot
        model, frames, auds = head
        from b2nerf.render import FramePipeline
        pipe = FramePipeline(model, N_RAYS, depth=6, torso=m, bg_coords=coords)
        for sl in pipe.slots:
            sl.set_torso_pose(pose)
        k = [0]

        def one():
            f = k[0] % len(frames); k[0] += 1
            pipe.submit_device(frames[f][0], frames[f][1], auds[f])

        for _ in range(12):
            one()
        pipe.drain(); torch.cuda.synchronize(); e0.record()
        for _ in range(60):
            one()
        pipe.drain(); e1.record(); torch.cuda.synchronize()
        frames_with_torso = 60 / (e0.elapsed_time(e1) * 1e-3)
    return {"pixels": HW * HW, "torso_pixels": n_on, "fused_ms_per_frame": fused_ms, "fused_eager_call_ms": fused_eager_ms, "per_op_ms_per_frame": per_op_ms,
            "fused_gflops": 2.0 * macs * n_on / (fused_ms * 1e-3) / 1e9, "frames_per_sec_head_plus_torso": frames_with_torso,
            "train_fwd_bwd_fused_ms": train_fused_ms, "train_fwd_bwd_per_op_autograd_ms": train_per_op_ms,
            "note": "k_torso_frame: occupancy test + freq encoding + deform MLP + tiled fp16 grid + torso MLP + blend in one launch (CUDA cores, thread = pixel); "
                    "per-op = run_torso on the drop-in encoders + torch Linear under autocast (~25 launches)"}


def reference_on_b200():
    """The number to beat: the unmodified reference model on its own CUDA extensions on this GPU, in a separate process (its packages shadow the
    drop-ins).  Needs oracle/_ref_py + oracle/_ref (staged / built in the build container, shipped with the snapshot)."""
    try:
        r = subprocess.run([sys.executable, os.path.join(ROOT, "profiles", "reference_on_b200.py"), "--frames", "20", "--steps", "20"], capture_output=True, text=True,
                           timeout=600)
        last = [ln for ln in r.stdout.strip().splitlines() if ln.startswith("{")]
        return json.loads(last[-1]) if last else {"unavailable": (r.stderr or r.stdout)[-300:]}
    except Exception as e:                                 # noqa: BLE001
        return {"unavailable": f"{type(e).__name__}: {e}"}


def run_gpu_arm(args, rank, world, local_rank):
    from b2nerf import lib
    from b2nerf.render import FramePipeline
    dev = torch.device("cuda", local_rank)
    torch.cuda.set_device(dev)
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=dev)
    frames_np, auds_np, bitfield = synthetic_inputs(rank)
    model = build_model(dev)
    model.density_bitfield.copy_(torch.from_numpy(bitfield).to(dev))
    from b2nerf import scene as _scene
    pipe = FramePipeline(model, N_RAYS, depth=1 if args.no_graph else max(1, args.in_flight), use_graph=not args.no_graph,
                         camera=(HW, HW) + tuple(_scene.intrinsics(HW, HW)))
    r = pipe.slots[0]
    frames = [(torch.from_numpy(o).to(dev), torch.from_numpy(d).to(dev)) for o, d in frames_np]
    auds = [torch.from_numpy(a).to(dev) for a in auds_np]
    host_o = [torch.from_numpy(o).pin_memory() for o, _ in frames_np]
    host_d = [torch.from_numpy(d).pin_memory() for _, d in frames_np]
    host_a = [torch.from_numpy(a).pin_memory() for a in auds_np]
    out_hosts = [torch.empty(N_RAYS, 3).pin_memory() for _ in range(pipe.depth)]
    out_host = out_hosts[0]
    L = lib()

    def barrier():
        if world > 1:
            import torch.distributed as dist
            dist.barrier()
        torch.cuda.synchronize(dev)

    def timed(fn, steps, warmup):
        """`fn(k)` submits frame k; a step is FRAMES_PER_STEP frames."""
        for s in range(warmup * FRAMES_PER_STEP):
            fn(s)
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        l0 = L.launch_count()
        e0.record()
        for s in range(steps * FRAMES_PER_STEP):
            fn(warmup * FRAMES_PER_STEP + s)
        pipe.drain()                     # the timing stream waits for every frame in flight
        e1.record()
        barrier()
        ms = e0.elapsed_time(e1)
        if world > 1:
            import torch.distributed as dist
            t = torch.tensor([ms], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms, L.launch_count() - l0

    sampler = ClockSampler(local_rank) if rank == 0 else None
    if sampler:
        sampler.start()
        sampler.wait_first()
    dev_ms, launches = timed(lambda s: pipe.submit_device(frames[s % POOL][0], frames[s % POOL][1], auds[s % POOL]), args.steps, args.warmup)
    e2e_ms, _ = timed(lambda s: pipe.submit_host(host_o[s % POOL], host_d[s % POOL], host_a[s % POOL], out_hosts[s % pipe.depth]), args.steps, args.warmup)
    clocks = sampler.stop() if sampler else None          # sampled over both timed regions (device-resident and end-to-end)
    # the lighter end-to-end variant: pose + audio window up, RGB24 frame down, rays generated / image packed on the device
    pose_ms = None
    if r.loop_graph is not None and r.fused_audio:
        host_p = [torch.from_numpy(_scene.camera_pose(rank * 1000 + f).astype(np.float32)).pin_memory() for f in range(POOL)]
        out_u8 = [torch.empty(N_RAYS, 3, dtype=torch.uint8).pin_memory() for _ in range(pipe.depth)]
        pose_ms, _ = timed(lambda s: pipe.submit_host_pose(host_p[s % POOL], host_a[s % POOL], out_u8[s % pipe.depth]), args.steps, args.warmup)
    train_info = None
    if not args.no_train:
        train_info = train_bench(dev, rank, world, bitfield, barrier, steps=max(64, args.steps * 32), warmup=max(args.warmup, 18), eager=args.train_eager,
                                 fused_head=not args.train_unfused, kernel_breakdown=(world == 1 and not args.no_kernels))
    if rank != 0:
        return
    torso_info = None
    if not args.no_train:                                  # a secondary leg: never lose the headline line over it
        try:
            torso_info = torso_bench(dev, head=(model, frames, auds))
        except Exception as e:                             # noqa: BLE001
            torso_info = {"error": f"{type(e).__name__}: {e}"}
    img = out_host.numpy()
    assert np.isfinite(img).all() and 0.0 < float(img.mean()) <= 1.0
    if r.loop_graph is not None:        # device-controlled WHILE loop: kernels per frame = fixed + per-iteration x iterations actually executed
        iters = r.last_iterations()
        gpu_launches = int((r.kernels_fixed + r.kernels_per_iteration * iters) * args.steps * FRAMES_PER_STEP)
        loop_mode = f"one CUDA graph per frame, WHILE conditional node, {iters} iterations executed in the last frame"
    elif r.graph is not None:
        gpu_launches, loop_mode = int(r.launches_per_frame * args.steps * FRAMES_PER_STEP), "torch CUDA graph of the fixed 16-iteration sequence"
    else:
        gpu_launches, loop_mode = int(launches), "eager launches"
    n_frames = args.steps * FRAMES_PER_STEP
    value = world * n_frames / (dev_ms * 1e-3)
    e2e = world * n_frames / (e2e_ms * 1e-3)
    e2e_rays = {"value": e2e, "unit": "frames/s", "h2d_bytes_per_step": FRAMES_PER_STEP * r.h2d_bytes(), "d2h_bytes_per_step": FRAMES_PER_STEP * r.d2h_bytes(),
                "ms_per_step": e2e_ms / args.steps,
                "api": "FramePipeline.submit_host: pinned rays_o / rays_d / audio window in, pinned float image out"}
    pk, pk_src = peaks()
    head_ms, head_launches, head_samples = head_kernel_profile(model, r, frames, auds, min(args.steps, 8))
    flops = 2.0 * MACS_PER_SAMPLE_INFER * head_samples
    achieved = flops / (head_ms * 1e-3) / 1e12
    peak = float(pk.get("bf16_tflops", pk.get("bf16_tflops_sustained")))          # burst peak: every head launch is timed in isolation
    line = {
        "metric": METRIC, "value": value, "unit": "frames/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": dev_ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f16", "data": "synthetic",
        "config": {"workload": "infer_512x512_frame", "rays_per_frame": N_RAYS, "max_steps": 16, "dt_gamma": 1 / 256, "bound": 1,
                   "params": 683509, "weights": "random-init", "frames_per_gpu_per_step": FRAMES_PER_STEP, "timed_region_s": dev_ms * 1e-3, "frames_in_flight_per_gpu": pipe.depth, "parallelism": f"frames sharded over {world} GPU(s), no collective",
                   "l2": f"inputs cycle through {POOL} distinct frames ({POOL * N_RAYS * 24 / 1e6:.0f} MB of rays) > 126 MB L2", "cuda_graph": not args.no_graph, "loop": loop_mode},
        # End to end through the public API with HOST buffers.  The reference's inference call takes a head pose + the audio window (provider_for_inference.py:597-605:
        # poses.to(device), get_rays ON the device) and hands (preds * 255).astype(uint8) bytes to its frame queue (TrainerUtil.py:668): that is the primary `e2e`
        # (FrameRenderer.render_host_pose: 64 B + the audio window up, one RGB24 frame down; rays and the uint8 packing on the device).  The heavier variant that ships
        # the 6.3 MB of precomputed rays up and the float image down is kept next to it — at 8 GPUs it is bound by the host's memory / PCIe root, not by the GPUs.
        "e2e": e2e_rays if pose_ms is None else {"value": world * n_frames / (pose_ms * 1e-3), "unit": "frames/s",
                                                  "h2d_bytes_per_step": FRAMES_PER_STEP * (64 + host_a[0].numel() * 4),
                                                  "d2h_bytes_per_step": FRAMES_PER_STEP * N_RAYS * 3, "ms_per_step": pose_ms / args.steps,
                                                  "api": "FramePipeline.submit_host_pose: pinned 4x4 pose + audio window in, pinned RGB24 frame out"},
        "e2e_rays_in_float_out": e2e_rays,
        "gpu_launches": gpu_launches,
        "clocks": clocks,
        "roofline": {"kernel": "k_head_infer4 (fused tri-plane gather + 7 tcgen05 layers, activations in tensor memory)", "bound": "tensor", "achieved": achieved, "peak": peak, "unit": "TFLOP/s",
                     "frac": achieved / peak, "traffic": ncu_traffic(), "peak_source": pk_src + " (bf16 burst: the launches are timed alone)",
                     "how": f"average launch duration from CUDA events on the launching stream (5 back-to-back repeats per launch) for each of the "
                            f"{head_launches} head launches of {min(args.steps, 8)} frames, separate un-graphed pass of the same frames; "
                            f"{head_samples} samples x {2 * MACS_PER_SAMPLE_INFER} FLOP",
                     "head_ms_per_frame": head_ms / min(args.steps, 8), "samples_per_frame": head_samples / min(args.steps, 8)},
    }
    if not args.no_train:
        line["train"] = train_info
        line["torso"] = torso_info
    if world == 1 and not args.no_kernels:
        try:                                               # per-kernel rooflines of the per-op kernels at sizes > L2 (profiles/kernel_rooflines.py)
            sys.path.insert(0, os.path.join(ROOT, "profiles"))
            import kernel_rooflines
            del model, pipe, frames
            torch.cuda.empty_cache()
            line["kernels"] = kernel_rooflines.measure()
        except Exception as e:                             # noqa: BLE001
            line["kernels"] = {"error": f"{type(e).__name__}: {e}"}
    if world == 1 and not args.no_ref_gpu:
        line["reference_on_b200"] = reference_on_b200()
    if world == 1 and not args.no_cpu:
        rows = HW // ROW_STRIDE                            # the reference arm's sample: every 8th row of the frame
        fps, sec, ns = cpu_frames_per_sec(rows, 2, 1)
        line["cpu_baseline"] = {"value": fps, "unit": "frames/s", "cores": os.cpu_count(), "kind": "port",
                                "sample": f"{rows} rows of one 512x512 frame at an even stride of {ROW_STRIDE} ({rows * HW} rays, {ns} samples, x{ROW_STRIDE} extrapolation), "
                                          f"mean of 2 after 1 warm-up, pure-PyTorch fp32 (oracle/torch_port.py)"}
    _emit(line)


_JSON_OUT = None


def _claim_stdout():
    """The contract is ONE JSON line on stdout.  Libraries write there too (NCCL prints its version banner on stdout when the first communicator is made), so
    file descriptor 1 is pointed at stderr for the whole run and the JSON line alone goes to the original stdout."""
    global _JSON_OUT
    if _JSON_OUT is None:
        sys.stdout.flush()
        _JSON_OUT = os.fdopen(os.dup(1), "w")
        os.dup2(2, 1)


def _emit(line):
    out = _JSON_OUT if _JSON_OUT is not None else sys.stdout
    out.write(json.dumps(line) + "\n")
    out.flush()


def main():
    _claim_stdout()
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b2nerf", choices=["b2nerf", "reference"])
    ap.add_argument("--no-graph", action="store_true")
    ap.add_argument("--in-flight", type=int, default=8, help="frames in flight per GPU (independent frames on separate streams; 1 = strictly one after another)")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--no-train", action="store_true", help="skip the training-step leg")
    ap.add_argument("--no-kernels", action="store_true", help="skip the per-kernel roofline leg (N = 1)")
    ap.add_argument("--no-ref-gpu", action="store_true", help="skip the reference-on-this-GPU leg (N = 1)")
    ap.add_argument("--train-eager", action="store_true", help="training leg without the CUDA graph")
    ap.add_argument("--train-unfused", action="store_true", help="training leg through torch autograd over the per-op kernels instead of the fused head kernels")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "b2nerf" else args.warmup
    rank, world, local_rank = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
    if args.impl == "reference":
        run_reference_arm(args, rank)
        return
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device — the product path has no CPU fallback (use --impl reference for the CPU arm)")
    run_gpu_arm(args, rank, world, local_rank)
    if world > 1:
        import torch.distributed as dist
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
