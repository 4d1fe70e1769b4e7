"""Run one of the reference's scripts with `fmc` bound to the B200 mirror, without editing the reference:

    python -m synfmc_b200.launch /path/to/SynFMC/train_cam_obj_ctrl.py --config configs/obj.yaml
    torchrun --nproc-per-node 8 -m synfmc_b200.launch /path/to/SynFMC/train_cam_obj_ctrl.py --config ...

The script's directory is taken as the reference checkout (FMC_REFERENCE_ROOT) so that the parts of `fmc` outside the
hot path (datasets) come from it; see synfmc_b200/dropin.py."""
import os
import runpy
import sys


def main(argv=None):
    argv = list(sys.argv[1:] if argv is None else argv)
    if not argv or argv[0] in ("-h", "--help"):
        print(__doc__)
        return 2
    script = os.path.abspath(argv[0])
    if not os.path.isfile(script):
        raise SystemExit(f"synfmc_b200.launch: no such script: {script}")
    from . import dropin
    root = os.path.dirname(script)
    dropin.install(reference_root=root if os.path.isdir(os.path.join(root, "fmc")) else None)
    sys.argv = [script] + argv[1:]
    runpy.run_path(script, run_name="__main__")
    return 0


if __name__ == "__main__":
    sys.exit(main())
