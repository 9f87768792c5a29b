"""The C-ABI library loads and exports every symbol include/dazim_b200.h declares (no compute)."""
import ctypes
import os
import re

from conftest import ROOT


def _declared():
    txt = open(os.path.join(ROOT, "include", "dazim_b200.h")).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    names = re.findall(r"\b([a-z_0-9]+)\s*\(", txt)
    return sorted({n for n in names if n.startswith("dazim_") or n.endswith("_")})


def test_exports_match_header():
    from dazimsurftomo_b200 import build
    lib = ctypes.CDLL(build.build())
    names = _declared()
    assert "dazim_gbuild" in names and "calsurfganisojoint_" in names and len(names) >= 20
    missing = [n for n in names if not hasattr(lib, n)]
    assert not missing, missing


def test_no_device_fails_loudly():
    import torch
    if torch.cuda.is_available():
        return
    from dazimsurftomo_b200 import api
    try:
        api.Handle(0)
    except api.DazimError as e:
        assert e.code >= 100   # DAZIM_ECUDA + cudaError: there is no CPU fallback
    else:
        raise AssertionError("dazim_create succeeded without a GPU")
