"""Drop-in contract: the mirror classes expose exactly the reference's state_dict keys/shapes
for the real 512 / 256 configs (fixture generated from the reference by oracle/make_golden.py)."""
import gzip
import json
import os

import pytest
import torch

from conftest import GOLDEN


@pytest.fixture(scope="module")
def keys():
    with gzip.open(os.path.join(GOLDEN, "state_dict_keys.json.gz"), "rt") as f:
        return json.load(f)


@pytest.mark.parametrize("tag", ["512", "256"])
def test_unet_keys(keys, tag):
    from dynamicrafter_b200.lvdm.modules.networks.openaimodel3d import UNetModel
    with torch.device("meta"):
        m = UNetModel(**keys["unet_cfg_" + tag])
    got = {k: list(v.shape) for k, v in m.state_dict().items()}
    assert got == keys["unet_" + tag]
    assert len(got) == (1516 if tag == "512" else 1532)


def test_decoder_and_autoencoder_keys(keys):
    from dynamicrafter_b200.lvdm.modules.networks.ae_modules import Decoder
    from dynamicrafter_b200.lvdm.models.autoencoder import AutoencoderKL
    with torch.device("meta"):
        d = Decoder(**keys["ddconfig_512"])
        ae = AutoencoderKL(ddconfig=keys["ddconfig_512"], embed_dim=4)
    assert {k: list(v.shape) for k, v in d.state_dict().items()} == keys["decoder_512"]
    got = {k: list(v.shape) for k, v in ae.state_dict().items()}
    assert got == keys["autoencoder_512"]


def test_zero_init_matches_reference_convention():
    """zero_module'd tensors are zero at construction like the reference (openaimodel3d.py:179,545)."""
    from dynamicrafter_b200.lvdm.modules.networks.openaimodel3d import UNetModel
    from oracle.make_golden import TINY_UNET
    m = UNetModel(**TINY_UNET)
    sd = m.state_dict()
    assert float(sd["out.2.weight"].abs().max()) == 0.0
    assert float(sd["input_blocks.1.0.out_layers.3.weight"].abs().max()) == 0.0
    assert float(sd["input_blocks.1.1.proj_out.weight"].abs().max()) == 0.0
    assert float(sd["input_blocks.1.0.temopral_conv.conv4.3.weight"].abs().max()) == 0.0
    assert float(sd["fps_embedding.2.weight"].abs().max()) == 0.0
