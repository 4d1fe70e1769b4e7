"""Make `import fmc...` resolve to this package's mirror (`synfmc_b200.fmc`) -- the level-1 integration of
INTEGRATION.md.

Putting `synfmc_b200/` on PYTHONPATH is not enough, for two reasons: (1) `python train_cam_ctrl.py` puts the
reference checkout (which has its own `fmc/`) at sys.path[0], ahead of PYTHONPATH; (2) the mirror's modules import
their siblings relatively (`from ... import ops`) and must therefore stay `synfmc_b200.fmc.*` modules.  So `install()`
registers a meta-path finder in FRONT of the path finder that answers every `fmc` / `fmc.*` import with the
corresponding `synfmc_b200.fmc.*` module OBJECT (one module, two names).

What the mirror does not provide -- the data-loading classes `UnrealTrajVideoDataset` / `UnrealTrajLoraDataset`
(fmc/data/dataset.py:979,2215) that the trainers import next to `ray_condition`, or whole modules outside the hot path
-- is served from the reference checkout itself when its root is known (`reference_root=` or FMC_REFERENCE_ROOT): its
`fmc/` directory is mounted as the private package `_fmc_reference`, so the reference's CPU code runs unmodified, with
its own relative imports, and nothing of it is copied here."""
import importlib
import importlib.abc
import importlib.util
import os
import sys
import types

MIRROR = "synfmc_b200.fmc"
REFERENCE_PACKAGE = "_fmc_reference"


def reference_root():
    root = os.environ.get("FMC_REFERENCE_ROOT")
    return root if root and os.path.isdir(os.path.join(root, "fmc")) else None


def reference_module(submodule):
    """import `fmc.<submodule>` of the reference checkout as `_fmc_reference.<submodule>` (None when no checkout is known)"""
    root = reference_root()
    if root is None:
        return None
    if REFERENCE_PACKAGE not in sys.modules:
        pkg = types.ModuleType(REFERENCE_PACKAGE)
        pkg.__path__ = [os.path.join(root, "fmc")]
        pkg.__package__ = REFERENCE_PACKAGE
        sys.modules[REFERENCE_PACKAGE] = pkg
    return importlib.import_module(f"{REFERENCE_PACKAGE}.{submodule}" if submodule else REFERENCE_PACKAGE)


def reference_attr(submodule, name):
    """module-level __getattr__ helper (PEP 562) for mirror modules that only carry the hot-path part of a reference module"""
    mod = reference_module(submodule)
    if mod is None:
        raise AttributeError(
            f"fmc.{submodule}.{name} is not part of the B200 mirror (only the denoising hot path is); set "
            f"FMC_REFERENCE_ROOT to a FudanCVL/SynFMC checkout (or use `python -m synfmc_b200.launch <script>`) and it is "
            f"taken from the reference's own fmc/{submodule.replace('.', '/')}.py")
    return getattr(mod, name)


class _FmcAlias(importlib.abc.MetaPathFinder, importlib.abc.Loader):
    def find_spec(self, fullname, path=None, target=None):
        if fullname != "fmc" and not fullname.startswith("fmc."):
            return None
        real = MIRROR + fullname[3:]
        try:
            found = importlib.util.find_spec(real)
        except ModuleNotFoundError:
            found = None
        if found is not None:
            return importlib.util.spec_from_loader(fullname, self, is_package=found.submodule_search_locations is not None)
        # a module the mirror does not have at all: the reference's own, if a checkout is known
        if reference_root() is not None:
            ref = f"{REFERENCE_PACKAGE}{fullname[3:]}"
            reference_module("")
            try:
                found = importlib.util.find_spec(ref)
            except ModuleNotFoundError:
                found = None
            if found is not None:
                return importlib.util.spec_from_loader(fullname, self, is_package=found.submodule_search_locations is not None)
        return None

    def create_module(self, spec):
        real = MIRROR + spec.name[3:]
        try:
            if importlib.util.find_spec(real) is not None:
                return importlib.import_module(real)
        except ModuleNotFoundError:
            pass
        return importlib.import_module(f"{REFERENCE_PACKAGE}{spec.name[3:]}")

    def exec_module(self, module):  # the aliased module is already executed
        return None


_finder = None


def install(reference_root=None):
    """Idempotent.  After this, `import fmc`, `from fmc.models.unet import ...` etc. bind the B200 mirror."""
    global _finder
    if reference_root is not None:
        os.environ["FMC_REFERENCE_ROOT"] = os.path.abspath(reference_root)
    if _finder is None:
        stale = [m for m in sys.modules if m == "fmc" or m.startswith("fmc.")]
        if stale:
            raise RuntimeError(f"synfmc_b200.dropin.install(): `fmc` is already imported from elsewhere ({stale[0]}); "
                               "install the alias before the first `import fmc`")
        _finder = _FmcAlias()
        sys.meta_path.insert(0, _finder)
    return _finder
