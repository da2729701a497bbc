"""GPU parity of the sampler paths beside plain sampling: the mask / x0 blend (ddim.py:174-180), q_sample
(ddpm3d.py:305-308), stochastic_encode / decode (ddim.py:282-319) -- against outputs of the unmodified reference
(tests/golden/ddim_mask.pt, made by oracle/make_golden.py mask)."""
import pytest
import torch

from conftest import load_golden, rel_l2
from dynamicrafter_b200.synth import synth_input

pytestmark = pytest.mark.gpu


def _setup(tag):
    from oracle import ddim_oracle
    from dynamicrafter_b200.lvdm.models.samplers.ddim import DDIMSampler
    from test_models_gpu import _MockDiffusion, SCHED
    c = SCHED[tag]
    gr = 0.7 if tag == "c2" else 0.0
    tab = ddim_oracle.DiffusionTables(**c["kw"])
    dev = torch.device("cuda")
    model = _MockDiffusion(tab, lambda x, t, cnd, **kw: torch.tanh(0.3 * x + cnd), dev)
    smp = DDIMSampler(model)
    smp.batch_cfg = False          # the mock conditioning is a broadcast scalar, not a per-sample tensor
    return smp, model, tab, c, gr, dev


@pytest.mark.parametrize("tag", ["c2", "c1"])
def test_mask_x0_blend_path_vs_reference_golden(tag):
    g = load_golden("ddim_mask.pt")
    smp, model, tab, c, gr, dev = _setup(tag)
    shape = (2, 4, 16, 8, 8)
    mask = (synth_input("mask", (2, 1, 16, 8, 8), 61) > 0).float().cuda()
    x0, x_T = synth_input("x0", shape, 61).cuda(), synth_input("x_T", shape, 61).cuda()
    y, _ = smp.sample(S=c["S"], batch_size=2, shape=shape[1:], conditioning=torch.tensor([0.1], device=dev),
                      unconditional_conditioning=torch.tensor([-0.1], device=dev), unconditional_guidance_scale=7.5,
                      eta=0.0, verbose=False, timestep_spacing=c["spacing"], guidance_rescale=gr, x_T=x_T, mask=mask,
                      x0=x0, clean_cond=True)
    err = rel_l2(y.cpu(), g[f"{tag}_clean1"])
    print(f"\nmask/x0 path {tag}: {c['S']}-step rel-L2 vs reference golden {err:.3e}")
    assert err < 1e-4          # tanh differs by ~1 ulp between CPU and GPU; compounds over the steps


@pytest.mark.parametrize("tag", ["c2", "c1"])
def test_stochastic_encode_bit_exact_and_decode(tag):
    g = load_golden("ddim_mask.pt")
    smp, model, tab, c, gr, dev = _setup(tag)
    shape = (2, 4, 16, 8, 8)
    x0, noise = synth_input("x0", shape, 61).cuda(), synth_input("enc_noise", shape, 61).cuda()
    smp.make_schedule(c["S"], ddim_discretize=c["spacing"], ddim_eta=0.0, verbose=False)
    t_enc = g[f"{tag}_t_enc"]
    z = smp.stochastic_encode(x0, torch.tensor([t_enc] * 2, device=dev), noise=noise)
    assert torch.equal(z.cpu(), g[f"{tag}_enc"])           # two rounded products + one rounded sum: bit-exact
    dec = smp.decode(z, torch.tensor([0.1], device=dev), t_enc, unconditional_guidance_scale=7.5,
                     unconditional_conditioning=torch.tensor([-0.1], device=dev))
    err = rel_l2(dec.cpu(), g[f"{tag}_dec"])
    print(f"\nstochastic_encode + decode {tag}: rel-L2 vs reference golden {err:.3e}")
    assert err < 1e-4


def test_q_sample_bit_exact_vs_torch():
    """host.q_sample (the non-clean_cond blend) against torch's eager expression on the same noise."""
    from dynamicrafter_b200 import ops
    x0 = synth_input("x0", (3, 4, 16, 8, 8), 5).cuda()
    nz = synth_input("nz", (3, 4, 16, 8, 8), 5).cuda()
    a = torch.tensor([0.9, 0.5, 0.01], device="cuda")
    c = torch.tensor([0.3, 0.8, 0.999], device="cuda")
    ref = a.view(3, 1, 1, 1, 1) * x0 + c.view(3, 1, 1, 1, 1) * nz
    assert torch.equal(ops.q_sample(x0, nz, a, c), ref)
