"""Host-side (Python / launch) cost of one end-to-end step: cProfile over the e2e loop of bench.py (pinned host inputs,
result read back every step).  python tools/prof_e2e_host.py"""
import cProfile, os, pstats, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from dynamicrafter_b200.configs import SAMPLER
from dynamicrafter_b200.lvdm.models.samplers.ddim import DDIMSampler

dev = torch.device("cuda", 0)
torch.manual_seed(bench.SEED)
host, mp = bench.build_host("512", dev, with_vae=False)
host.unet.use_cuda_graph = True
sm = SAMPLER["512"]
B, T = 1, 16
pin = bench.make_inputs(mp, B, T, dev, pinned=True)
fs = torch.tensor([sm["fs"]] * B, dtype=torch.long, device=dev)
smp = DDIMSampler(host)
smp.make_schedule(sm["S"], ddim_discretize=sm["timestep_spacing"], ddim_eta=sm["eta"], verbose=False)
steps_desc = list(reversed(list(smp.ddim_timesteps)))
total = len(steps_desc)
hx = pin["x_T"].clone().pin_memory()
out_host = torch.empty_like(hx).pin_memory()


def step(i):
    dx = hx.to(dev, non_blocking=True)
    c_i = {"c_crossattn": [pin["cond"].to(dev, non_blocking=True)], "c_concat": [pin["concat"].to(dev, non_blocking=True)]}
    uc_i = {"c_crossattn": [pin["uncond"].to(dev, non_blocking=True)], "c_concat": c_i["c_concat"]}
    st = int(steps_desc[i % total])
    ts = torch.full((B,), st, device=dev, dtype=torch.long)
    smp._step_hint = st
    xo = smp.p_sample_ddim(dx, c_i, ts, index=total - (i % total) - 1, unconditional_guidance_scale=sm["cfg"],
                           unconditional_conditioning=uc_i, fs=fs, guidance_rescale=sm["guidance_rescale"])[0]
    out_host.copy_(xo, non_blocking=True)
    t_enq = time.perf_counter()
    torch.cuda.synchronize()
    hx.copy_(out_host)
    return t_enq


for i in range(5):
    step(i)
N = 30
t0 = time.perf_counter()
enq = 0.0
for i in range(N):
    ts_ = time.perf_counter()
    enq += step(i) - ts_
print(f"e2e step {(time.perf_counter() - t0) / N * 1e3:.3f} ms, of which host enqueue (until the last launch returns) {enq / N * 1e3:.3f} ms")
pr = cProfile.Profile()
pr.enable()
for i in range(N):
    step(i)
pr.disable()
pstats.Stats(pr).sort_stats("cumulative").print_stats(14)
