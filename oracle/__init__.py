"""CPU oracle for the DynamiCrafter denoising hot path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``dynamicrafter_b200/`` may import this
package; only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s
``cpu_baseline`` / ``--impl reference`` legs do, and only as the checker / the
reported CPU baseline.

The oracle is a *functional restatement* (plain torch fp32 on CPU, no nn.Module,
weights read straight out of a ``state_dict``) of the reference algorithm:

* ``unet_oracle.unet_forward``      <- lvdm/modules/networks/openaimodel3d.py:548-603
* ``vae_oracle.decode``             <- lvdm/models/autoencoder.py:104-107,
                                       lvdm/modules/networks/ae_modules.py:539-578
* ``ddim_oracle.sample``            <- lvdm/models/samplers/ddim.py:24-279 and
                                       lvdm/models/samplers/ddim_multiplecond.py:220-236 (three guidance branches)
* ``vae_oracle.encode``             <- lvdm/models/autoencoder.py:97-102, ae_modules.py:438-463
* ``resampler_oracle.resampler_forward`` <- lvdm/modules/encoders/resampler.py:48-145
* ``output_oracle.video_to_uint8``  <- scripts/evaluation/inference.py:153-159
* ``sds_oracle.sds_target`` (+ ``sample_timestep``, ``add_noise``, ``apply_guidance_rescale``)
                                    <- guidance_pipeline.py:273-424 (SDS guidance core)

Parity pin: the reference ships no tests or golden vectors, so
the oracle is pinned against outputs of the reference itself, imported in the
build container from /root/reference by ``oracle/make_golden.py`` and committed
as small fixtures under ``tests/golden/``.  ``tests/test_oracle_golden.py``
re-checks the oracle against those fixtures on every run (CPU).
"""
