"""Branch-sharded sampling (cond on rank 0, uncond on rank 1, predictions all-gathered every step) vs the same sampler
in one process with the CFG pair as one batch.  Run under torchrun with 2 ranks (see frame_shard_check.py)."""
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from conftest import load_golden, rel_l2                                             # noqa: E402
from dynamicrafter_b200.synth import synth_state_dict, synth_input                   # noqa: E402


def main():
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    local = int(os.environ.get("LOCAL_RANK", rank))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    os.environ["NCCL_DEBUG"] = "WARN"
    dist.init_process_group("nccl", device_id=dev)
    from oracle import ddim_oracle
    from dynamicrafter_b200.lvdm.modules.networks.openaimodel3d import UNetModel
    from dynamicrafter_b200.lvdm.models.samplers.ddim import DDIMSampler
    from dynamicrafter_b200.parallel import BranchShard
    from test_models_gpu import _MockDiffusion, SCHED
    g = load_golden("ddim.pt")["e2e_c2"]
    ucfg, seed = g["ucfg"], g["seed"]
    unet = UNetModel(**ucfg)
    unet.load_state_dict(synth_state_dict(g["unet_shapes"], seed), strict=True)
    unet = unet.cuda().eval()
    tab = ddim_oracle.DiffusionTables(**SCHED["c2"]["kw"])
    B, T, h, w = 1, 16, 8, 8
    cond = {"c_crossattn": [synth_input("cond", (B, 77 + 16 * T, 1024), seed).cuda()],
            "c_concat": [synth_input("concat", (B, 4, T, h, w), seed).cuda()]}
    uc = {"c_crossattn": [synth_input("uncond", (B, 77 + 16 * T, 1024), seed).cuda()], "c_concat": cond["c_concat"]}
    x_T = synth_input("x_T", (B, 4, T, h, w), seed).cuda()
    fs = torch.tensor([ucfg["default_fs"]] * B, dtype=torch.long, device=dev)

    def apply_model(x, t, cnd, **kw):
        return unet(torch.cat([x] + cnd["c_concat"], dim=1), t, context=torch.cat(cnd["c_crossattn"], 1), **kw)

    def run(shard):
        smp = DDIMSampler(_MockDiffusion(tab, apply_model, dev))
        smp.branch_shard = shard
        torch.manual_seed(11)
        out, _ = smp.sample(S=g["S"], batch_size=B, shape=(4, T, h, w), conditioning=cond, unconditional_guidance_scale=7.5,
                            unconditional_conditioning=uc, eta=1.0, verbose=False, x_T=x_T, fs=fs,
                            timestep_spacing=g["spacing"], guidance_rescale=g["gr"])
        return out

    one = run(None)
    two = run(BranchShard())
    torch.cuda.synchronize()
    same = [torch.empty_like(two) for _ in range(world)]
    dist.all_gather(same, two)
    identical = all(torch.equal(s, same[0]) for s in same)
    err = rel_l2(two.cpu(), one.cpu())
    if rank == 0:
        print(f"branch-sharded vs batched CFG, {g['S']} steps eta=1: rel-L2 {err:.3e}, ranks identical: {identical}", flush=True)
    ok = err < 0.1 and identical          # two bf16 evaluations (different GroupNorm reduction order), CFG-amplified
    flag = torch.tensor([1.0 if ok else 0.0], device=dev)
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    dist.destroy_process_group()
    sys.exit(0 if float(flag) == 1.0 else 1)


if __name__ == "__main__":
    main()
