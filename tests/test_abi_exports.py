"""The C-ABI library loads without a GPU and exports every entry point include/vlr_engine.h declares; entry points
that need a device fail loudly (no CPU fallback) when there is none."""
import os
import re

import pytest

from varlociraptor_b200 import contamination as ct, engine

HEADER = os.path.join(os.path.dirname(__file__), "..", "include", "vlr_engine.h")


def declared_functions():
    text = re.sub(r"/\*.*?\*/", "", open(HEADER).read(), flags=re.S)
    return sorted(set(re.findall(r"^\s*(?:const\s+)?[A-Za-z_][\w\s\*]*?\b(vlr_\w+)\s*\(", text, flags=re.M)))


def test_every_declared_entry_point_is_exported_and_bound():
    names = declared_functions()
    assert "vlr_call_batch" in names and "vlr_contamination_posterior" in names and len(names) >= 15
    lib = engine.lib()
    for name in names:
        assert hasattr(lib, name), "libvlr_engine.so does not export %s" % name
    assert sorted(engine.EXPORTED_SYMBOLS) == names


def test_contamination_model_needs_a_device():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(engine.EngineError, match="no CUDA device"):
        ct.contamination_posterior([], None, device=0)
