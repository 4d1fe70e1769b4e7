"""Drop-in mirror of the reference's `fmc` package surface for the denoising hot path: same module,
class, argument and state-dict key names as FudanCVL/SynFMC `fmc/`, executed by hand-written sm_100a kernels through
libfmc_b200.so.  Put `synfmc_b200/` on sys.path (or import `synfmc_b200.fmc`) and the trainers' imports
`from fmc.models.unet import UNet3DConditionModelPoseCond` etc. resolve here."""
