"""Hyper-parameters of the three released DynamiCrafter models (values of configs/inference_{256,512,1024}_v1.0.yaml
in the reference; 1024 differs from 512 only in image_size, base_scale and default_fs)."""
import copy

_UNET = dict(in_channels=8, out_channels=4, model_channels=320, attention_resolutions=[4, 2, 1], num_res_blocks=2,
             channel_mult=[1, 2, 4, 4], dropout=0.1, num_head_channels=64, transformer_depth=1, context_dim=1024,
             use_linear=True, use_checkpoint=False, temporal_conv=True, temporal_attention=True,
             temporal_selfatt_only=True, use_relative_position=False, use_causal_attention=False, temporal_length=16,
             addition_attention=True, image_cross_attention=True, fs_condition=True)
_VAE = dict(embed_dim=4, ddconfig=dict(double_z=True, z_channels=4, resolution=256, in_channels=3, out_ch=3, ch=128,
                                       ch_mult=[1, 2, 4, 4], num_res_blocks=2, attn_resolutions=[], dropout=0.0))
_COMMON = dict(linear_start=0.00085, linear_end=0.012, timesteps=1000, scale_factor=0.18215)

MODEL_PARAMS = {
    "256": dict(_COMMON, parameterization="eps", image_size=[32, 32],
                unet_config=dict(params=dict(_UNET, image_cross_attention_scale_learnable=True, default_fs=3)),
                first_stage_config=dict(params=_VAE)),
    "512": dict(_COMMON, parameterization="v", rescale_betas_zero_snr=True, use_dynamic_rescale=True, base_scale=0.7,
                perframe_ae=True, image_size=[40, 64],
                unet_config=dict(params=dict(_UNET, default_fs=24)), first_stage_config=dict(params=_VAE)),
    "1024": dict(_COMMON, parameterization="v", rescale_betas_zero_snr=True, use_dynamic_rescale=True, base_scale=0.3,
                 perframe_ae=True, image_size=[72, 128],
                 unet_config=dict(params=dict(_UNET, default_fs=10)), first_stage_config=dict(params=_VAE)),
}

# generative frame interpolation / looping video (configs[4] of BASELINE.json; scripts/run_application.sh): the 512 model
# with two-keyframe conditioning -- c_concat holds the first and last frame latents and zeros in between
# (inference.py:245-248) -- and fs = 5
MODEL_PARAMS["512_interp"] = MODEL_PARAMS["512"]

# sampler settings of scripts/run.sh in the reference
SAMPLER = {
    "256": dict(S=50, eta=1.0, cfg=7.5, timestep_spacing="uniform", guidance_rescale=0.0, fs=3),
    "512": dict(S=50, eta=1.0, cfg=7.5, timestep_spacing="uniform_trailing", guidance_rescale=0.7, fs=24),
    "1024": dict(S=50, eta=1.0, cfg=7.5, timestep_spacing="uniform_trailing", guidance_rescale=0.7, fs=10),
    "512_interp": dict(S=50, eta=1.0, cfg=7.5, timestep_spacing="uniform_trailing", guidance_rescale=0.7, fs=5,
                       keyframes_only=True),
}

# algorithmic FLOPs (2*MAC, conv + matmul only) of ONE UNet forward (B=1, 16 frames) and of the
# 16-frame VAE decode, in TFLOP
UNET_TFLOP = {"256": 4.904, "512": 12.603, "1024": 52.336, "512_interp": 12.603}
VAE_TFLOP = {"256": 9.955, "512": 25.016, "1024": 92.069, "512_interp": 25.016}


def model_params(name):
    return copy.deepcopy(MODEL_PARAMS[name])
