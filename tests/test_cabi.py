"""The C-ABI library loads without a GPU and exports every symbol include/fmc_b200.h declares (no compute calls)."""
import os
import re

from synfmc_b200 import _cabi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    text = open(os.path.join(ROOT, "include", "fmc_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(fmc_[a-z0-9_]+)\s*\(", text)))


def test_header_matches_binding_table():
    declared = set(_declared())
    bound = set(_cabi.SIGNATURES) | {"fmc_abi_version", "fmc_last_error_string", "fmc_groupnorm_launches",
                                     "fmc_grad_norm_workspace_floats", "fmc_colsum_workspace_floats",
                                     "fmc_layernorm_bwd_blocks", "fmc_groupnorm_bwd_workspace_floats", "fmc_wgrad_workspace_floats"}
    assert declared == bound, (declared - bound, bound - declared)


def test_library_exports_every_declared_symbol():
    handle = _cabi.lib()
    for name in _declared():
        assert hasattr(handle, name), name
    assert handle.fmc_abi_version() == _cabi.ABI_VERSION


def test_signature_arity_matches_header():
    text = open(os.path.join(ROOT, "include", "fmc_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    for name, argtypes in _cabi.SIGNATURES.items():
        m = re.search(r"\b" + name + r"\s*\((.*?)\)\s*;", text, flags=re.S)
        assert m, name
        n_args = len([a for a in m.group(1).split(",") if a.strip()])
        assert n_args == len(argtypes), (name, n_args, len(argtypes))


def test_ops_refuse_cpu_tensors():
    """The product path must fail loudly instead of falling back when there is no CUDA tensor."""
    import pytest
    import torch

    from synfmc_b200 import ops
    with pytest.raises(_cabi.FmcError):
        ops.gemm(torch.zeros(8, 8, dtype=torch.bfloat16), torch.zeros(8, 8, dtype=torch.bfloat16))
    from oracle import harness as helpers
    from synfmc_b200.fmc.models.pose_adaptor import CameraPoseEncoder
    enc = CameraPoseEncoder(channels=[320, 640], **helpers.POSE_ENCODER_KWARGS)
    with pytest.raises(RuntimeError):
        enc(torch.zeros(1, 6, 2, 64, 64))
