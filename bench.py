#!/usr/bin/env python
"""Benchmark of the DynamiCrafter denoising hot path on B200 (contract: see the task statement).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--config 512]

One "step" = one classifier-free-guided DDIM step of the 320x512x16-frame model (BASELINE.json configs[1]):
cond + uncond UNet forward (issued as one batch-2 forward) + the fused CFG/DDIM update.  `value` = steps/s with
all inputs resident in HBM; `e2e` = the same step driven through the public DDIMSampler API with the step's inputs
(x_t, conditioning) coming from pinned HOST memory and x_{t-1} read back to the host every step.  With N>1 GPUs
every rank denoises its own video (the reference's own scaling mode: independent prompts per rank) -> weak scaling,
no collective on the data path; the final latents are all-gathered once (NCCL) outside the timed steps.

Extra objects: `roofline` (tcgen05 implicit-GEMM kernel: algorithmic FLOPs / CUDA-event time of every launch of one
step, against the measured bf16 peak), `cpu_baseline` (the CPU oracle restatement of the reference, timed on the
host cores on a bounded sample), `breakdown_ms` (per kernel class), `video` (full 50-step sample + VAE decode).
"""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

METRIC = "UNet denoise steps/s at 320x512x16f, 50 DDIM, CFG (1 step = cond+uncond UNet forward + DDIM update)"
SEED = 7


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return dict(tflops=d.get("bf16_tflops_sustained", 1379.1), burst=d.get("bf16_tflops", 1636.3),
                    hbm=d.get("hbm_gbs", 6535.4), source="measured (MEASURED_PEAKS.json, sustained)")
    return dict(tflops=1400.0, burst=1590.0, hbm=6650.0, source="fallback (B200_PROFILING.md)")


def _hbm_entry(v, replay, peaks):
    """v = [algorithmic bytes, per-launch-event ms] of a bandwidth-bound kernel class; replay = ms of the class re-issued as
    one graph (None if not measured)."""
    ms = replay if isinstance(replay, float) else v[1]
    gbs = v[0] / (ms * 1e-3) / 1e9
    out = {"achieved_gbs": round(gbs, 1), "peak_gbs": peaks["hbm"], "frac": round(gbs / peaks["hbm"], 4), "ms_per_step": round(ms, 3)}
    if isinstance(replay, float):
        out["frac_per_launch_events"] = round(v[0] / (v[1] * 1e-3) / 1e9 / peaks["hbm"], 4)
    return out


def load_traffic():
    """roofline.traffic: DRAM bytes per launch of the dominant kernel from the committed ncu captures (null if absent)."""
    p = os.path.join(ROOT, "profiles", "r02_traffic.json")
    if not os.path.exists(p):
        return {"traffic": None}
    d = json.load(open(p))
    w, c = d["gemm_res"], d["gemm_conv"]
    return {"traffic": w["traffic"], "traffic_algorithmic": w["traffic_algorithmic"],
            "traffic_unit": "bytes/launch, " + w["what"] + " (" + w["capture"] + ")",
            "traffic_other_shape": {"what": c["what"], "traffic": c["traffic"], "traffic_algorithmic": c["traffic_algorithmic"],
                                    "capture": c["capture"]}}


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 50 ms while the timed region runs."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.proc, self.path = index, None, None

    def start(self):
        try:
            fd, self.path = tempfile.mkstemp(suffix=".csv")
            os.close(fd)
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "50"],
                                         stdout=open(self.path, "w"), stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons, pw = [], [], set(), []
        for line in open(self.path):
            f = [s.strip() for s in line.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0])); mx.append(float(f[1])); pw.append(float(f[2]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        os.unlink(self.path)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        busy = sorted(sm)[len(sm) // 2:]
        return {"sm_mhz": sorted(busy)[len(busy) // 2], "sm_max_mhz": max(mx), "reasons": sorted(reasons),
                "power_w_max": max(pw), "samples": len(sm)}


def _graph_collectives():
    from dynamicrafter_b200.lvdm import _packing
    return _packing.GRAPH_COLLECTIVES


def build_host(cfg_name, device, with_vae=True):
    from dynamicrafter_b200.configs import model_params
    from dynamicrafter_b200.host import DiffusionHost
    from dynamicrafter_b200.synth import synth_tensor
    mp = model_params(cfg_name)
    if not with_vae:
        mp.pop("first_stage_config")
    with torch.device("meta"):
        host = DiffusionHost.from_model_params(mp)
    host = host.to_empty(device=device)
    # to_empty() discards buffer values: rebuild the schedule buffers from a CPU instance without networks
    ref = DiffusionHost.from_model_params(dict(mp, unet_config=dict(params=_tiny_unet()), first_stage_config=None))
    for name, buf in ref.named_buffers():
        if "." not in name:
            getattr(host, name).copy_(buf)
    with torch.no_grad():
        for mod, prefix in ((host.unet, "unet."), (host.first_stage_model, "vae.")):
            if mod is None:
                continue
            for k, p in list(mod.named_parameters()) + list(mod.named_buffers()):
                p.copy_(synth_tensor(prefix + k, p.shape, SEED).to(device))
    return host.eval(), mp


def _tiny_unet():
    return dict(in_channels=8, out_channels=4, model_channels=64, attention_resolutions=[1], num_res_blocks=1,
                channel_mult=[1], num_head_channels=64, context_dim=64, use_linear=True, temporal_length=16,
                use_relative_position=False)


def make_inputs(mp, B, T, device, pinned=False):
    from dynamicrafter_b200.synth import synth_input
    h, w = mp["image_size"]
    t = {"x_T": synth_input("x_T", (B, 4, T, h, w), SEED), "concat": synth_input("concat", (B, 4, T, h, w), SEED),
         "cond": synth_input("cond", (B, 77 + 16 * T, 1024), SEED), "uncond": synth_input("uncond", (B, 77 + 16 * T, 1024), SEED)}
    if pinned:
        return {k: v.pin_memory() for k, v in t.items()}
    return {k: v.to(device) for k, v in t.items()}


# ------------------------------------------------------------------------------------------- instrumentation
class OpTimer:
    """Wraps the ops.* launch helpers with CUDA events (on the launching stream) for ONE eager step."""
    CLASSES = {"tap_gemm": "tcgen05_gemm_conv", "flash_attention": "attention_spatial",
               "temporal_attention": "attention_temporal", "group_norm": "groupnorm", "layer_norm": "layernorm",
               "ddim_update": "ddim_update"}

    def __init__(self, retain=False):
        """retain=True: keep every launch's arguments alive for replay_class (a separate pass: holding a step's tensors makes
        the allocator cudaMalloc, which would distort the per-launch events of the timed pass)."""
        self.retain = retain
        from dynamicrafter_b200 import ops
        self.ops, self.saved, self.records, self.gemm = ops, {}, [], []
        self.calls = {}                          # class -> [(name, args, kwargs)] for replay_class

    def __enter__(self):
        ops = self.ops
        names = ["tap_gemm", "flash_attention", "temporal_attention", "group_norm", "layer_norm", "ddim_update",
                 "softmax_rows", "nchw_to_rows", "rows_to_nchw", "upsample2x", "space_to_depth", "cast_bf16",
                 "timestep_embedding", "small_linear"]
        for n in names:
            fn = getattr(ops, n)
            self.saved[n] = fn
            setattr(ops, n, self._wrap(n, fn))
        return self

    def _wrap(self, name, fn):
        def inner(*a, **kw):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            out = fn(*a, **kw)
            e1.record()
            flops, key, nbytes = 0, name, 0
            if name == "tap_gemm":
                dims, taps, w, N = a[1], a[4], a[5], a[6]
                kdim = w.shape[1] // (3 if a[0].dtype == torch.float32 else 1)     # fp32 mode: algorithmic K, not 3K
                flops = 2.0 * dims[0] * dims[1] * dims[2] * N * kdim
                key = (f"gemm M={dims[0] * dims[1] * dims[2]} N={N} K={kdim} taps={len(taps)} "
                       f"f32={kw.get('out_f32') is not None} bf16={kw.get('out_bf16') is not None} "
                       f"res={kw.get('residual') is not None} geglu={bool(kw.get('geglu'))}")
            elif name == "group_norm":
                key = f"group_norm G={a[1]} S={a[2]} C={a[0].shape[1] + (kw['x2'].shape[1] if kw.get('x2') is not None else 0)}"
                # algorithmic bytes: every source element read once, every output element written once
                outs = out if isinstance(out, tuple) else (out,)
                nbytes = a[0].numel() * a[0].element_size() + sum(o.numel() * o.element_size() for o in outs)
                if kw.get("x2") is not None:
                    nbytes += kw["x2"].numel() * kw["x2"].element_size()
            elif name == "flash_attention":
                key = f"flash_attention B={a[2]} H={a[3]} nq={a[4]} nk={[s[2] for s in a[1]]}"
            elif name == "layer_norm":
                key = f"layer_norm rows={a[0].shape[0]} C={a[0].shape[1]}"
                nbytes = a[0].numel() * a[0].element_size() + out.numel() * out.element_size()
            self.records.append((self.CLASSES.get(name, "layout_misc"), e0, e1, flops, key, nbytes))
            if self.retain and name in ("tap_gemm", "group_norm", "layer_norm"):
                self.calls.setdefault(self.CLASSES[name], []).append((name, a, kw))
            return out
        return inner

    def __exit__(self, *exc):
        for n, fn in self.saved.items():
            setattr(self.ops, n, fn)

    def replay_class(self, cls, reps=3):
        """Kernel time of one class without the per-launch event overhead: every launch of the class in the step, with the
        step's own operands (all still alive; together far larger than L2), re-issued back to back as ONE CUDA graph and
        timed with CUDA events around the replay.  Programmatic dependent launch overlaps the prologues as in the real
        step graph.  Returns ms per step's worth of launches."""
        calls = self.calls.get(cls, [])
        if not calls:
            return None

        def issue():
            for name, a, kw in calls:
                self.saved[name](*a, **kw)
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            issue()
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            issue()
        g.replay()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            g.replay()
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / reps
        del g
        return ms

    def summary(self):
        torch.cuda.synchronize()
        by, gemm_ms, gemm_flops, n_gemm, shapes = {}, 0.0, 0.0, 0, {}
        self.hbm = {}                              # class -> [algorithmic bytes, ms] of the bandwidth-bound kernels
        for cls, e0, e1, fl, key, nb in self.records:
            ms = e0.elapsed_time(e1)
            by[cls] = by.get(cls, 0.0) + ms
            if nb:
                h = self.hbm.setdefault(cls, [0, 0.0])
                h[0] += nb; h[1] += ms
            ent = shapes.setdefault(key, [0, 0.0, 0.0])
            ent[0] += 1; ent[1] += ms; ent[2] += fl
            if cls == "tcgen05_gemm_conv":
                gemm_ms += ms; gemm_flops += fl; n_gemm += 1
        dump = os.environ.get("DCB_BENCH_SHAPES")
        if dump:       # per-shape table for kernel tuning: count, total ms, TFLOP/s
            rows = sorted(shapes.items(), key=lambda kv: -kv[1][1])
            with open(dump, "w") as f:
                for k, (n, ms, fl) in rows:
                    f.write(f"{ms:9.3f} ms  x{n:<4d} {fl / max(ms, 1e-9) / 1e9:8.1f} TFLOP/s  {k}\n")
        return by, gemm_ms, gemm_flops, n_gemm, len(self.records)


# ------------------------------------------------------------------------------------------- CPU baseline
def cpu_sample_seconds(cfg_name, frames, threads):
    """One cond-branch UNet forward of the CPU oracle (fp32 eager restatement of the reference) on `frames` of the
    16 frames of the workload latent.  Returns (seconds, description)."""
    from dynamicrafter_b200.configs import model_params
    from dynamicrafter_b200.host import DiffusionHost
    from dynamicrafter_b200.synth import synth_input
    from oracle import unet_oracle
    torch.set_num_threads(threads)
    mp = model_params(cfg_name)
    ucfg = mp["unet_config"]["params"]
    with torch.device("meta"):
        shapes = {k: tuple(v.shape) for k, v in DiffusionHost(ucfg).unet.state_dict().items()}
    g = torch.Generator().manual_seed(SEED)
    sd = {}
    for k, s in shapes.items():      # cheap non-degenerate init: timing does not depend on the values
        t = torch.empty(s)
        t.uniform_(-1, 1, generator=g)
        if len(s) >= 2:
            t.mul_(1.0 / (t[0].numel() ** 0.5))
        elif k.endswith("weight"):
            t.mul_(0.1).add_(1.0)
        sd[k] = t
    h, w = mp["image_size"]
    x = synth_input("x", (1, 8, frames, h, w), SEED)
    ctx = synth_input("ctx", (1, 77 + 16 * frames, 1024), SEED)
    ts = torch.tensor([500], dtype=torch.long)
    fs = torch.tensor([24], dtype=torch.long)
    return sd, (lambda: unet_oracle.unet_forward(sd, ucfg, x, ts, ctx, fs))


def _reference_available():
    from oracle import ref_host
    return ref_host.ref_root() is not None


def workload_name(config):
    """config.workload, identical in both arms (the reference arm times the same step of the same workload)."""
    from dynamicrafter_b200.configs import SAMPLER, model_params
    sm = SAMPLER[config]
    h, w = model_params(config)["image_size"]
    return (f"DynamiCrafter {config} ({h * 8}x{w * 8}x16f, latent {h}x{w}), DDIM {sm['S']} {sm['timestep_spacing']} eta "
            f"{sm['eta']}, CFG {sm['cfg']}, guidance_rescale {sm['guidance_rescale']}; 1 video per GPU; random-init weights "
            "(1.44 B params)")


class _StdoutToStderr:
    """The reference prints while it builds ("AE working on z of shape ..."); stdout must carry the ONE JSON line only."""

    def __enter__(self):
        sys.stdout.flush()
        self.saved = os.dup(1)
        os.dup2(2, 1)

    def __exit__(self, *exc):
        sys.stdout.flush()
        os.dup2(self.saved, 1)
        os.close(self.saved)


def time_reference_steps(config, steps, warmup, budget_s):
    with _StdoutToStderr():
        return _time_reference_steps(config, steps, warmup, budget_s)


def _time_reference_steps(config, steps, warmup, budget_s):
    """Times the UNMODIFIED reference (oracle/_ref, staged by oracle/build_ref.py) on the host cores through its own
    public API: one step = DDIMSampler.p_sample_ddim at the full workload shape (cond forward + uncond forward +
    update; no frame or branch extrapolation).  K is honoured up to a wall-clock budget; returns the per-step times
    actually measured."""
    from oracle import ref_host
    yaml_name = {"256": "inference_256_v1.0.yaml", "512": "inference_512_v1.0.yaml", "1024": "inference_1024_v1.0.yaml",
                 "512_interp": "inference_512_v1.0.yaml"}[config]
    t_build = time.perf_counter()
    step, info = ref_host.reference_step_runner(yaml_name, threads=os.cpu_count(), seed=SEED)
    info["build_s"] = time.perf_counter() - t_build
    t_start = time.perf_counter()
    times = []
    for i in range(warmup + steps):
        t0 = time.perf_counter()
        step()
        dt = time.perf_counter() - t0
        if i >= warmup:
            times.append(dt)
        elapsed = time.perf_counter() - t_start
        # stop early once another step would overrun the budget (always keep >= 1 timed step)
        if i + 1 >= warmup + 1 and elapsed + dt > budget_s and i + 1 < warmup + steps:
            if i < warmup:                       # budget exhausted inside the warm-up: time exactly one more step
                t0 = time.perf_counter()
                step()
                times.append(time.perf_counter() - t0)
            break
    return times, info


def run_reference(args):
    """--impl reference: the reference's own CPU implementation of the path on the box's host cores.  oracle/_ref
    (the unmodified reference, pure Python, staged by oracle/build_ref.py) when present -> kind "reference"; else the
    oracle port on a bounded sample -> kind "port"."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    threads = os.cpu_count()
    budget = float(os.environ.get("DCB_REF_BUDGET_S", "150"))
    if _reference_available():
        warm = min(args.warmup, 1)
        times, info = time_reference_steps(args.config, args.steps, warm, budget)
        step_s = sum(times) / len(times)
        n_steps = len(times)
        kind = "reference"
        h, w = info["latent"]
        sample = (f"{n_steps} full DDIM step(s) of the unmodified reference ({info['api']}), latent {h}x{w}x{info['frames']}f, "
                  f"2 UNet forwards per step, {threads} threads, {warm} warm-up step(s); no extrapolation; "
                  f"--steps {args.steps} requested, bounded by a {budget:.0f} s budget")
    else:
        frames = 2
        sd, fwd = cpu_sample_seconds(args.config, frames, threads)
        for _ in range(min(args.warmup, 1)):
            fwd()
        times = []
        for _ in range(args.steps):
            t0 = time.perf_counter()
            fwd()
            times.append(time.perf_counter() - t0)
        step_s = sum(times) / len(times) * (16 / frames) * 2                  # 16 frames, cond + uncond
        n_steps, warm, kind = args.steps, min(args.warmup, 1), "port"
        sample = (f"oracle UNet forward (fp32 eager torch, {threads} threads), 1 CFG branch, {frames} of 16 frames at latent "
                  f"40x64; step time extrapolated x{16 // frames} frames x2 branches (oracle/_ref not staged)")
    val = 1.0 / step_s
    line = {"impl": "reference", "metric": METRIC, "value": val, "unit": "steps/s", "n_gpus": args.gpus, "steps": n_steps,
            "warmup": warm, "steps_requested": args.steps, "ms_per_step": step_s * 1e3, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": workload_name(args.config), "videos_per_gpu": 1,
                       "parallelism": "host CPU threads (the reference has no GPU-free path other than torch CPU)",
                       "note": "reference default-init weights; timing does not depend on the values"},
            "cpu_baseline": {"value": val, "unit": "steps/s", "cores": threads, "kind": kind, "sample": sample},
            "e2e": {"value": val, "unit": "steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)



# ------------------------------------------------------------------------------------------- N > 1 extras
def run_shard_check(dev, rank, world):
    """Frame-sharded forward of a tiny UNet (every temporal block behind an all-to-all, 5-D GroupNorm sums all-reduced)
    against the same forward on one GPU, in the fp32 parity mode where no bf16 rounding noise hides a wrong transpose:
    {rel_l2, identical (all ranks hold the same bits), frames}."""
    import torch.distributed as dist
    from dynamicrafter_b200.lvdm.modules.networks.openaimodel3d import UNetModel
    from dynamicrafter_b200.parallel import FrameShard
    from dynamicrafter_b200.precision import use_precision
    from dynamicrafter_b200.synth import synth_state_dict, synth_input
    cfg = dict(in_channels=8, out_channels=4, model_channels=64, attention_resolutions=[1, 2], num_res_blocks=1,
               channel_mult=[1, 2], num_head_channels=64, context_dim=64, use_linear=True, temporal_length=16,
               use_relative_position=False, temporal_conv=True, temporal_attention=True, image_cross_attention=True,
               fs_condition=True, default_fs=10)
    T, h, w = 16, 8, 8
    m = UNetModel(**cfg)
    m.load_state_dict(synth_state_dict({k: tuple(v.shape) for k, v in m.state_dict().items()}, SEED))
    m = m.to(dev).eval()
    x = synth_input("x", (2, 8, T, h, w), SEED).to(dev)
    ctx = synth_input("context", (2, 77 + 16 * T, 64), SEED).to(dev)
    ts = torch.tensor([500, 493], dtype=torch.long, device=dev)
    res = {}
    for mode in ("fp32", "bf16"):
        with use_precision(mode):
            m.frame_shard = None
            y1 = m(x, ts, context=ctx)
            m.frame_shard = FrameShard()
            yN = m(x, ts, context=ctx)
            torch.cuda.synchronize()
        e = float((yN.double() - y1.double()).norm() / y1.double().norm())
        same = [torch.empty_like(yN) for _ in range(world)]
        dist.all_gather(same, yN)
        res[mode] = {"rel_l2_vs_single_gpu": e, "identical_on_all_ranks": all(torch.equal(t, same[0]) for t in same)}
    m.frame_shard = None
    res["frames_per_rank"] = T // world
    res["pass"] = bool(res["fp32"]["rel_l2_vs_single_gpu"] < 1e-4 and res["fp32"]["identical_on_all_ranks"]
                       and res["bf16"]["rel_l2_vs_single_gpu"] < 2.5e-2 and res["bf16"]["identical_on_all_ranks"])
    return res


def run_strong(host, dev, rank, world, args):
    """ONE 576x1024 video (BASELINE.json configs[3], latent 72x128) over all GPUs: frames sharded over the ranks (x the
    two guidance branches from 4 GPUs on), against the same step on one GPU.  UNet + update only, eager sharded mode."""
    import torch.distributed as dist
    from dynamicrafter_b200.configs import SAMPLER
    from dynamicrafter_b200.lvdm.models.samplers.ddim import DDIMSampler
    from dynamicrafter_b200.parallel import FrameShard, BranchShard
    from dynamicrafter_b200.synth import synth_input
    B, T, h, w = 1, 16, 72, 128
    sm = SAMPLER["1024"]
    x0 = synth_input("x_T", (B, 4, T, h, w), SEED).to(dev)
    concat = synth_input("concat", (B, 4, T, h, w), SEED).to(dev)
    cond = {"c_crossattn": [synth_input("cond", (B, 77 + 16 * T, 1024), SEED).to(dev)], "c_concat": [concat]}
    uc = {"c_crossattn": [synth_input("uncond", (B, 77 + 16 * T, 1024), SEED).to(dev)], "c_concat": [concat]}
    fs = torch.tensor([sm["fs"]] * B, dtype=torch.long, device=dev)

    def timed(smp, n_warm=2, n=3):
        smp.make_schedule(sm["S"], ddim_discretize=sm["timestep_spacing"], ddim_eta=sm["eta"], verbose=False)
        steps = list(reversed(list(smp.ddim_timesteps)))

        def one(x, i):
            ts = torch.full((B,), int(steps[i]), device=dev, dtype=torch.long)
            return smp.p_sample_ddim(x, cond, ts, index=len(steps) - i - 1, unconditional_guidance_scale=sm["cfg"],
                                     unconditional_conditioning=uc, fs=fs, guidance_rescale=sm["guidance_rescale"])[0]
        torch.manual_seed(SEED)
        x = x0.clone()
        for i in range(n_warm):
            x = one(x, i)
        dist.barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(n):
            x = one(x, n_warm + i)
        e1.record()
        dist.barrier()
        torch.cuda.synchronize()
        t = torch.tensor([e0.elapsed_time(e1) / n], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        assert torch.isfinite(x).all()
        return float(t)

    graph = host.unet.use_cuda_graph
    ms1 = timed(DDIMSampler(host))                               # every rank runs the whole video on its own GPU
    host.unet.__dict__.pop("_graphs", None)                      # free the 1024-shaped single-GPU graph
    torch.cuda.empty_cache()
    nf, fgroup, bgroup = world, None, None
    if world >= 4 and world % 2 == 0:                            # rank r: guidance branch r % 2, frame shard r // 2
        nf = world // 2
        for br in range(2):
            g = dist.new_group([r for r in range(world) if r % 2 == br])
            if rank % 2 == br:
                fgroup = g
        for fi in range(nf):
            g = dist.new_group([2 * fi, 2 * fi + 1])
            if rank // 2 == fi:
                bgroup = g
    smp = DDIMSampler(host)
    host.unet.frame_shard = FrameShard(fgroup)
    if bgroup is not None:
        smp.branch_shard = BranchShard(bgroup)
    from dynamicrafter_b200.parallel import LINK
    LINK.reset()
    try:
        msN = timed(smp)
    finally:
        host.unet.frame_shard = None
        host.unet.use_cuda_graph = graph
    # bytes handed to peers per step (host-side count of the eager collectives of the 2 + 3 steps just run), whole box
    lt = torch.tensor([LINK.bytes / 5.0, LINK.calls / 5.0], device=dev, dtype=torch.float64)
    dist.all_reduce(lt)
    # every temporal block moves its bf16 operand to the pixel side and its bf16 delta back:
    # (11, 10, 10, 8) blocks at (320, 640, 1280, 1280) channels and (1, 1/4, 1/16, 1/64) of the 72x128 pixels
    alg = sum(nb * 2 * T * h * w * 2 * c // d for nb, c, d in ((11, 320, 1), (10, 640, 4), (10, 1280, 16), (8, 1280, 64)))
    nvlink = {"sent_to_peers_gb_per_step": round(float(lt[0]) / 1e9, 3), "collectives_per_rank_per_step": round(float(lt[1]) / world, 1),
              "algorithmic_gb_per_step": round(2 * alg / 1e9, 3),
              "note": "both guidance branches; algorithmic = SURVEY 8(e): whole bf16 tensors of the 39 temporal blocks, there "
                      "and back (3.6 GB per branch); sent = off-rank blocks only ((N-1)/N of it) + statistics + gathers"}

    def graph_fn():
        """The same sharded step with each rank's forward replayed as ONE CUDA graph that holds its NCCL kernels.  Run last
        and under a watchdog (`_finish_with_graph_strong`): such graphs cannot be torn down cleanly (torch 2.11 + NCCL 2.28)."""
        from dynamicrafter_b200.lvdm import _packing
        _packing.GRAPH_COLLECTIVES = True
        host.unet.__dict__.pop("_graphs", None)
        host.unet.frame_shard = FrameShard(fgroup)
        host.unet.use_cuda_graph = True
        ms = timed(smp, n_warm=3, n=5)
        return {"ms_per_step": round(ms, 2), "speedup": round(ms1 / ms, 3),
                "timing": "CUDA events, max over ranks, 3 warm-up + 5 timed steps; collectives captured in the graph"}

    return {"workload": "ONE 576x1024x16f video (latent 72x128), CFG DDIM step, UNet + update",
            "mode": f"frames sharded x{nf}" + (" x 2 guidance branches" if bgroup is not None else ""),
            "ms_per_step_1gpu": round(ms1, 2), "ms_per_step": round(msN, 2), "speedup": round(ms1 / msN, 3),
            "n_gpus": world, "scaling": "strong", "timing": "CUDA events, max over ranks, 2 warm-up + 3 timed steps",
            "nvlink": nvlink, "_graph_fn": graph_fn}


def _finish_with_graph_strong(line, strong, graph_fn, rank, limit_s=75.0):
    """Last act of an N > 1 run: measure the graph-captured sharded step, print THE line, leave with os._exit.  A watchdog
    prints the line without that number and ends the process if the attempt has not returned within `limit_s` (a hung
    collective cannot be caught as an exception), so the headline numbers never depend on this auxiliary measurement."""
    import threading
    import torch.distributed as dist
    lock, state = threading.Lock(), {"printed": False}

    def emit(res):
        with lock:
            if state["printed"]:
                return
            state["printed"] = True
            if rank == 0:
                strong["graph"] = res
                print(json.dumps(line), flush=True)

    def leave():
        sys.stdout.flush()
        sys.stderr.flush()
        os._exit(0)

    def bail():
        emit({"unavailable": f"no result within {limit_s:.0f} s"})
        leave()

    dog = threading.Timer(limit_s, bail)
    dog.daemon = True
    dog.start()
    try:
        res = graph_fn()
    except Exception as exc:
        res = {"unavailable": f"{type(exc).__name__}: {exc}"}
    emit(res)
    dog.cancel()
    dog = threading.Timer(30.0, leave)
    dog.daemon = True
    dog.start()
    try:
        dist.barrier()
        torch.cuda.synchronize()
    except Exception:
        pass
    leave()


# ------------------------------------------------------------------------------------------- main arm
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", default="512", choices=["256", "512", "1024", "512_interp"])
    ap.add_argument("--no-graph", action="store_true", help="launch kernels eagerly instead of replaying a CUDA graph")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-video", action="store_true")
    ap.add_argument("--precision", default="bf16", choices=["bf16", "fp32"],
                    help="fp32 = the <=1e-3 parity mode (hi/lo-split operands, 3x GEMM work, CUDA-core attention); "
                         "the headline numbers are the bf16 mode")
    ap.add_argument("--branch-shard", action="store_true",
                    help="N=2: ONE video, cond branch on rank 0 and uncond on rank 1, predictions all-gathered every step "
                         "(strong scaling of single-video latency)")
    ap.add_argument("--no-strong", action="store_true", help="N>1: skip the sharded-vs-single check and the strong-scaling "
                    "(one 576x1024 video over all GPUs) measurement appended to the line")
    ap.add_argument("--frame-shard", action="store_true",
                    help="N>1: ONE video, its 16 frames sharded over the ranks (all-to-all around the temporal blocks); "
                         "strong scaling of single-video latency (BASELINE.json configs[3])")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    if args.impl == "reference":
        return run_reference(args)

    import torch.distributed as dist
    from dynamicrafter_b200 import _lib, ops
    from dynamicrafter_b200.precision import set_precision
    set_precision(args.precision)
    from dynamicrafter_b200.configs import SAMPLER, UNET_TFLOP, VAE_TFLOP
    from dynamicrafter_b200.lvdm.models.samplers.ddim import DDIMSampler

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        # NCCL prints its version banner on stdout when the first communicator comes up: keep stdout to the single
        # JSON line by pointing fd 1 at stderr until that has happened
        sys.stdout.flush()
        saved = os.dup(1)
        os.dup2(2, 1)
        try:
            dist.init_process_group("nccl", device_id=dev)
            warm = torch.zeros(1, device=dev)
            dist.all_reduce(warm)
            torch.cuda.synchronize()
        finally:
            sys.stdout.flush()
            os.dup2(saved, 1)
            os.close(saved)
    _lib.check(_lib.lib().dcb_check_device(), "dcb_check_device")
    peaks = load_peaks()
    torch.manual_seed(SEED)                  # same per-step noise on every rank (sharded single-video modes)

    host, mp = build_host(args.config, dev, with_vae=not args.no_video)
    host.unet.use_cuda_graph = not args.no_graph
    fshard = args.frame_shard and world > 1
    bshard = args.branch_shard and world > 1
    fgroup = bgroup = None
    nf = world                                # ranks sharing the frames of one guidance branch
    if fshard and bshard:
        # 2 x (world / 2): rank r runs guidance branch r % 2 on frame shard r // 2
        assert world % 2 == 0 and world >= 4, "--frame-shard --branch-shard needs 4 or 8 ranks"
        nf = world // 2
        for br in range(2):
            g = dist.new_group([r for r in range(world) if r % 2 == br])
            if rank % 2 == br:
                fgroup = g
        for fi in range(nf):
            g = dist.new_group([2 * fi, 2 * fi + 1])
            if rank // 2 == fi:
                bgroup = g
    if fshard:
        from dynamicrafter_b200.parallel import FrameShard
        host.unet.frame_shard = FrameShard(fgroup)
    units = 1 if (fshard or bshard) else world          # videos in flight over the whole job
    sm = SAMPLER[args.config]
    B, T = 1, 16
    h, w = mp["image_size"]
    inp = make_inputs(mp, B, T, dev)
    pin = make_inputs(mp, B, T, dev, pinned=True)
    if sm.get("keyframes_only"):            # interpolation / looping: only the two key frames condition the video
        for d in (inp, pin):
            d["concat"][:, :, 1:-1] = 0
    fs = torch.tensor([sm["fs"]] * B, dtype=torch.long, device=dev)
    cond = {"c_crossattn": [inp["cond"]], "c_concat": [inp["concat"]]}
    uc = {"c_crossattn": [inp["uncond"]], "c_concat": [inp["concat"]]}
    smp = DDIMSampler(host)
    if bshard:
        from dynamicrafter_b200.parallel import BranchShard
        smp.branch_shard = BranchShard(bgroup)
    smp.make_schedule(sm["S"], ddim_discretize=sm["timestep_spacing"], ddim_eta=sm["eta"], verbose=False)
    steps_desc = list(reversed(list(smp.ddim_timesteps)))
    total = len(steps_desc)
    # DCB_BENCH_CFG: tuning-only override (1.0 = single-branch forward, to compare per-shape times at half the rows)
    kw = dict(unconditional_guidance_scale=float(os.environ.get("DCB_BENCH_CFG", sm["cfg"])),
              unconditional_conditioning=uc, fs=fs,
              guidance_rescale=sm["guidance_rescale"])

    def one_step(x, i, c=cond, extra=kw):
        i = i % total
        step = int(steps_desc[i])
        ts = torch.full((B,), step, device=dev, dtype=torch.long)
        smp._step_hint = step
        return smp.p_sample_ddim(x, c, ts, index=total - i - 1, **extra)[0]

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---------------- device-resident timing
    x = inp["x_T"].clone()
    for i in range(args.warmup):
        x = one_step(x, i)
    clocks = ClockSampler(local)
    barrier()
    clocks.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(args.steps):
        x = one_step(x, args.warmup + i)
    e1.record()
    barrier()
    ms = e0.elapsed_time(e1)
    clk = clocks.stop()
    t_ms = torch.tensor([ms], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t_ms, op=dist.ReduceOp.MAX)
    ms_per_step = float(t_ms) / args.steps
    value = units * B / (ms_per_step * 1e-3)
    assert torch.isfinite(x).all(), "non-finite latents"

    # ---------------- end-to-end through the public API from HOST buffers (pinned), result read back every step
    hx = pin["x_T"].clone().pin_memory()
    out_host = torch.empty_like(hx).pin_memory()
    h2d = sum(pin[k].numel() * 4 for k in ("x_T", "cond", "uncond", "concat"))
    d2h = hx.numel() * 4
    barrier()
    t0 = time.perf_counter()
    for i in range(args.steps):
        dx = hx.to(dev, non_blocking=True)
        c_i = {"c_crossattn": [pin["cond"].to(dev, non_blocking=True)], "c_concat": [pin["concat"].to(dev, non_blocking=True)]}
        uc_i = {"c_crossattn": [pin["uncond"].to(dev, non_blocking=True)], "c_concat": c_i["c_concat"]}
        xo = one_step(dx, args.warmup + i, c=c_i, extra=dict(kw, unconditional_conditioning=uc_i))
        out_host.copy_(xo, non_blocking=True)
        torch.cuda.synchronize()
        hx.copy_(out_host)
    barrier()
    e2e_ms = torch.tensor([(time.perf_counter() - t0) * 1e3], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(e2e_ms, op=dist.ReduceOp.MAX)
    e2e_value = units * B / (float(e2e_ms) / args.steps * 1e-3)

    # ---------------- instrumented eager step: per-launch CUDA events -> roofline of the tcgen05 kernel + breakdown
    host.unet.use_cuda_graph = False
    one_step(inp["x_T"].clone(), 0)
    torch.cuda.synchronize()
    calls0 = _lib.CALLS[0]
    with OpTimer() as ot:
        one_step(inp["x_T"].clone(), 1)
    by, gemm_ms, gemm_flops, n_gemm, n_launch = ot.summary()
    launches_per_step = _lib.CALLS[0] - calls0
    host.unet.use_cuda_graph = not args.no_graph
    achieved_ev = gemm_flops / (gemm_ms * 1e-3) / 1e12 if gemm_ms > 0 else 0.0
    # the same launches back to back as one CUDA graph (no per-launch events between them): the kernel's own time
    replay_ms = {}
    if not (fshard or bshard):
        try:
            host.unet.use_cuda_graph = False
            with OpTimer(retain=True) as ot2:
                one_step(inp["x_T"].clone(), 1)
            torch.cuda.synchronize()
            for cls in ("tcgen05_gemm_conv", "groupnorm", "layernorm"):
                replay_ms[cls] = ot2.replay_class(cls)
            ot2.calls = {}                       # release the step's tensors
            del ot2
            torch.cuda.empty_cache()
        except Exception as exc:
            replay_ms = {"error": f"{type(exc).__name__}: {exc}"}
        # the replayed GEMMs added into the step's (already consumed) statistics accumulators: clear them again
        ops.zero_stats_arena(dev)
        torch.cuda.synchronize()
        host.unet.use_cuda_graph = not args.no_graph
    gemm_replay = replay_ms.get("tcgen05_gemm_conv") if isinstance(replay_ms.get("tcgen05_gemm_conv"), float) else None
    achieved = gemm_flops / (gemm_replay * 1e-3) / 1e12 if gemm_replay else achieved_ev
    step_tflop = 2 * UNET_TFLOP[args.config] * B / (world if (fshard or bshard) else 1)      # per-GPU share of one step
    roofline = {"bound": "tensor", "kernel": "tap_gemm_kernel (tcgen05 implicit GEMM: linear + conv3x3 + temporal conv)",
                "achieved": round(achieved, 1), "peak": peaks["tflops"], "unit": "TFLOP/s",
                "frac": round(achieved / peaks["tflops"], 4),
                # dram__bytes_read.sum + dram__bytes_write.sum of ONE launch of this kernel, from the committed
                # `ncu --set full` captures of the current code (profiles/r02_traffic.json: the worst shape of the
                # step -- the HBM-bound 81920x320x320 + fp32 residual -- and a typical 3x3 conv)
                **load_traffic(),
                "peak_source": peaks["source"],
                "timing": ("all launches of one step re-issued back to back as one CUDA graph, CUDA events around the replay"
                           if gemm_replay else "per-launch CUDA events of one eager step"),
                "achieved_per_launch_events": round(achieved_ev, 1),
                "frac_per_launch_events": round(achieved_ev / peaks["tflops"], 4),
                "launches_per_step": n_gemm, "avg_launch_us": round((gemm_replay or gemm_ms) * 1e3 / max(n_gemm, 1), 2),
                "kernel_share_of_step": round((gemm_replay or gemm_ms) / ms_per_step, 4) if gemm_replay
                else round(gemm_ms / max(sum(by.values()), 1e-9), 4),
                "whole_step_tflops": round(step_tflop / (ms_per_step * 1e-3), 1),
                "whole_step_frac_of_peak": round(step_tflop / (ms_per_step * 1e-3) / peaks["tflops"], 4),
                "whole_step_frac_of_spec_2250": round(step_tflop / (ms_per_step * 1e-3) / 2250.0, 4),
                # the bandwidth-bound kernel classes of the same step: algorithmic bytes (each source element read
                # once, each output written once) / CUDA-event time, against the measured copy bandwidth
                "hbm_bound": {k: _hbm_entry(v, replay_ms.get(k), peaks) for k, v in sorted(ot.hbm.items())}}

    # ---------------- one full video through the public API: sample(50 steps) + all-gather + VAE decode
    video = None
    if not args.no_video:
        # the UNet has run its warm-up steps; give the decoder one too (lazy weight packing, allocator growth)
        host.decode_first_stage(inp["x_T"][:, :, :8].contiguous(), frames_per_call=8)
        barrier()
        t0 = time.perf_counter()
        samples, _ = smp.sample(S=sm["S"], batch_size=B, shape=(4, T, h, w), conditioning=cond,
                                unconditional_guidance_scale=sm["cfg"], unconditional_conditioning=uc, eta=sm["eta"],
                                verbose=False, x_T=inp["x_T"], fs=fs, timestep_spacing=sm["timestep_spacing"],
                                guidance_rescale=sm["guidance_rescale"])
        torch.cuda.synchronize()
        t_s = time.perf_counter() - t0
        from dynamicrafter_b200.parallel import all_gather_latents
        if fshard or bshard:                          # every rank holds the whole video: decode own frames only
            from dynamicrafter_b200.parallel import shard_range
            f0, f1 = shard_range(T, rank, world)
            t1 = time.perf_counter()
            frames = host.decode_first_stage(samples[:, :, f0:f1].contiguous(), frames_per_call=8)
        else:
            gathered = all_gather_latents(samples)        # the single collective of this mode
            assert gathered.shape[0] == world * B
            t1 = time.perf_counter()
            frames = host.decode_first_stage(samples, frames_per_call=8)
        torch.cuda.synchronize()
        t_d = time.perf_counter() - t1
        # VAE decode on its own: CUDA events around 3 repeated warm decodes of this rank's frames
        dz = (samples[:, :, f0:f1] if (fshard or bshard) else samples).contiguous()
        de0, de1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        de0.record()
        for _ in range(3):
            host.decode_first_stage(dz, frames_per_call=8)
        de1.record()
        torch.cuda.synchronize()
        dec_ms = de0.elapsed_time(de1) / 3
        dec_tflop = VAE_TFLOP[args.config] * dz.shape[0] * dz.shape[2] / 16.0
        tv = torch.tensor([t_s + t_d, t_s, t_d], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(tv, op=dist.ReduceOp.MAX)
        vid_tflop = 100 * UNET_TFLOP[args.config] + VAE_TFLOP[args.config]
        video = {"s_per_video": round(float(tv[0]), 3), "sample_s": round(float(tv[1]), 3), "decode_s": round(float(tv[2]), 3),
                 "videos_per_s_all_gpus": round(units / float(tv[0]), 4), "tflops": round(vid_tflop / float(tv[0]), 1),
                 "frac_of_peak": round(vid_tflop / float(tv[0]) / peaks["tflops"] / (world if (fshard or bshard) else 1), 4),
                 "finite": bool(torch.isfinite(frames).all()),
                 "vae_decode": {"ms": round(dec_ms, 2), "frames": int(dz.shape[0] * dz.shape[2]), "tflops": round(dec_tflop / (dec_ms * 1e-3), 1),
                                "frac_of_peak": round(dec_tflop / (dec_ms * 1e-3) / peaks["tflops"], 4),
                                "timing": "CUDA events, mean of 3 warm decodes (conv + GroupNorm + d=512 attention)"}}

    # ---------------- N > 1: sharded-vs-single check on a tiny UNet and the strong-scaling (single video) numbers
    shard_check = strong = None
    if world > 1 and not (fshard or bshard) and not args.no_strong:
        try:
            shard_check = run_shard_check(dev, rank, world)
            strong = run_strong(host, dev, rank, world, args)
        except Exception as exc:                     # never lose the headline line to the auxiliary measurements
            strong = {"error": f"{type(exc).__name__}: {exc}"}

    # ---------------- CPU baseline (rank 0, N=1): the unmodified reference (oracle/_ref) on a bounded sample
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        threads = os.cpu_count()
        if _reference_available():
            times, info = time_reference_steps(args.config, 1, 0, 1e9)
            cpu = {"value": 1.0 / times[0], "unit": "steps/s", "cores": threads, "kind": "reference",
                   "sample": f"1 full DDIM step of the unmodified reference ({info['api']}) at the workload shape "
                             f"(latent {h}x{w}x16f, cond + uncond forward + update) = {times[0]:.1f} s, un-warmed; "
                             "no extrapolation"}
        else:
            frames_s = 2
            _, fwd = cpu_sample_seconds(args.config, frames_s, threads)
            t0 = time.perf_counter()
            fwd()
            sec = time.perf_counter() - t0
            cpu = {"value": 1.0 / (sec * (16 / frames_s) * 2), "unit": "steps/s", "cores": threads, "kind": "port",
                   "sample": f"1 oracle UNet forward (1 CFG branch, {frames_s} of 16 frames, latent {h}x{w}) = {sec:.1f} s; "
                             f"x{16 // frames_s} frames x2 branches extrapolated"}

    if rank == 0:
        line = {"metric": METRIC, "value": round(value, 4), "unit": "steps/s", "n_gpus": world, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": round(ms_per_step, 3), "higher_is_better": True,
                "scaling": "strong" if (fshard or bshard) else "weak",
                "vs_baseline": None,
                "dtype": "bf16" if args.precision == "bf16" else "fp32 (bf16 hi/lo split x3 on tensor cores)",
                "data": "synthetic",
                "config": {"workload": workload_name(args.config),
                           "videos_per_gpu": B, "cuda_graph": (not args.no_graph) and (not fshard or _graph_collectives()),
                           "parallelism": (f"frame-sharded x{nf}" + (" x 2 guidance branches" if bshard else "") +
                                           f": one video, {T // nf} frames per GPU, all-to-all "
                                           "around the temporal blocks") if fshard else
                                          ((f"branch-sharded x{world}: one video, one guidance branch per GPU, predictions "
                                            "all-gathered every step") if bshard else f"one video per GPU x{world}"),
                           "l2": "not flushed: per-step working set (2.9 GB bf16 weights + activations) >> 126 MB L2"},
                "s_per_video_unet_only": round(sm["S"] * ms_per_step * 1e-3, 3),
                "e2e": {"value": round(e2e_value, 4), "unit": "steps/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h},
                "gpu_launches": int(launches_per_step * args.steps), "launches_per_step": int(launches_per_step),
                "clocks": clk, "roofline": roofline, "breakdown_ms": {k: round(v, 3) for k, v in sorted(by.items())},
                "video": video, "cpu_baseline": cpu}
        if shard_check is not None:
            line["shard_check"] = shard_check
        if strong is not None:
            line["strong"] = strong
    graph_fn = strong.pop("_graph_fn", None) if isinstance(strong, dict) else None
    if graph_fn is not None and not args.no_graph and os.environ.get("DCB_BENCH_STRONG_GRAPH", "1") != "0":
        _finish_with_graph_strong(line if rank == 0 else None, strong, graph_fn, rank)       # does not return
    if rank == 0:
        print(json.dumps(line), flush=True)
    if world > 1:
        from dynamicrafter_b200.lvdm import _packing
        if fshard and _packing.GRAPH_COLLECTIVES and not args.no_graph:
            # CUDA graphs that contain NCCL kernels: process-group / graph teardown hangs with torch 2.11 + NCCL 2.28,
            # so leave without it once every rank has finished (the JSON line is already out)
            dist.barrier()
            torch.cuda.synchronize()
            sys.stdout.flush()
            sys.stderr.flush()
            os._exit(0)
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
