"""synfmc_b200: B200-native implementation of the FMC (FudanCVL/SynFMC) pose-conditioned video-diffusion denoising
step.  `synfmc_b200.fmc` mirrors the reference's `fmc.models` / `fmc.pipelines` surface; `synfmc_b200.ops` are the
tensor-level entry points of libfmc_b200.so (include/fmc_b200.h)."""
__all__ = ["fmc", "ops", "engine", "synth", "dropin"]
