"""Full-size VAE decoder (ch 128, mult 1-2-4-4) on real latent sizes: warm decode throughput for 16 frames (parity
of a 320x512 frame against the oracle is tests/test_models_gpu.py::test_vae_decode_full_size_vs_oracle).
python tools/check_vae_full.py [h w]"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from dynamicrafter_b200.configs import model_params
from dynamicrafter_b200.lvdm.models.autoencoder import AutoencoderKL
from dynamicrafter_b200.synth import synth_state_dict, synth_input

h, w = (40, 64) if len(sys.argv) < 3 else (int(sys.argv[1]), int(sys.argv[2]))
fsc = model_params("512")["first_stage_config"]["params"]
ae = AutoencoderKL(ddconfig=fsc["ddconfig"], embed_dim=fsc["embed_dim"])
shapes = {k: tuple(v.shape) for k, v in ae.state_dict().items()}
sd = synth_state_dict(shapes, 11)
ae.load_state_dict(sd)
ae = ae.cuda().eval()
z = synth_input("z", (16, 4, h, w), 11) * 4.0
zc = z.cuda()
for _ in range(2):
    out = torch.cat([ae.decode(zc[i:i + 8]) for i in (0, 8)], 0)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
out = torch.cat([ae.decode(zc[i:i + 8]) for i in (0, 8)], 0)
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1)
tf = 25.016 * (h * w) / (40 * 64)
print(f"16 frames warm: {ms:.1f} ms  ({tf / ms * 1e3:.0f} TFLOP/s), finite {bool(torch.isfinite(out).all())}")
