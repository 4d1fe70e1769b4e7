"""The pipeline edges of SURVEY 8 row f3: B200-native stand-ins for the two third-party models the reference pipelines
call around the denoising loop -- diffusers' `AutoencoderKL` (decode_latents, fmc/pipelines/pipeline_animation.py:465-478;
vae.encode, train_cam_ctrl.py:544) and transformers' `CLIPTextModel` (_encode_prompt, pipeline_animation.py:480-567;
train_cam_ctrl.py:557-561).  Same constructor arguments, state-dict keys and call surface as the originals; executed by the
kernels of libfmc_b200.so (no eager fallback)."""
from .autoencoder_kl import AutoencoderKL  # noqa: F401
from .clip_text import CLIPTextModel  # noqa: F401
