"""Drop-in mirror of the reference's `fmc` package surface for the denoising hot path: same module,
class, argument and state-dict key names as FudanCVL/SynFMC `fmc/`, executed by hand-written sm_100a kernels through
libfmc_b200.so.  Import it as `synfmc_b200.fmc`, or call `synfmc_b200.dropin.install()` / run a reference script through
`python -m synfmc_b200.launch <script>` and the trainers' own imports (`from fmc.models.unet import
UNet3DConditionModelPoseCond` etc.) resolve here."""
