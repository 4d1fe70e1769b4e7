"""Call signatures of the mirror package against the reference's (tests/golden/fmc_reference_surface.json, recorded from
the reference's own classes and functions): every reference parameter must exist in the mirror at the same position,
with the same name, kind and default; the mirror may only APPEND optional parameters ("same names, argument
meaning")."""
import importlib
import inspect
import json
import os

import pytest

from tests.golden.make_golden_surface import SYMBOLS, describe, resolve

SURFACE = json.load(open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "fmc_reference_surface.json")))


def _norm(params):
    """The mirror never uses mutable defaults: a reference default of `{}` is `None` here (treated as empty), and list
    defaults are tuples.  Everything else must be equal."""
    out = []
    for name, default, kind in params:
        if default is not None:
            if default == "{}":
                default = "None"
            elif default.startswith("[") and default.endswith("]"):
                default = "(" + default[1:-1] + ")"
        out.append([name, default, kind])
    return out


@pytest.mark.parametrize("mod,qual", SYMBOLS, ids=[f"{m}:{q}" for m, q in SYMBOLS])
def test_signature_is_a_superset_of_the_reference(mod, qual):
    want = _norm(SURFACE[f"{mod}:{qual}"])
    got = _norm(describe(resolve(importlib.import_module("synfmc_b200.fmc." + mod), qual)))
    var_kw = [p for p in want if p[2] == "VAR_KEYWORD"]
    fixed = [p for p in want if p[2] not in ("VAR_KEYWORD", "VAR_POSITIONAL")]
    assert got[:len(fixed)] == fixed, f"\nreference: {fixed}\nmirror:    {got[:len(fixed)]}"
    extra = got[len(fixed):]
    for name, default, kind in extra:
        assert default is not None or kind in ("VAR_KEYWORD", "VAR_POSITIONAL"), f"extra required parameter {name}"
    if var_kw:
        assert any(k == "VAR_KEYWORD" for _, _, k in got), "the reference accepts **kwargs here"
