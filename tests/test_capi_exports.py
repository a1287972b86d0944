"""The C-ABI library loads without a GPU and exports every symbol include/mogp_b200.h declares; the ctypes
binding covers the same set.  (No compute calls here.)"""
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_functions():
    text = open(os.path.join(ROOT, "include", "mogp_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(mogp_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    from mogp_b200 import _capi
    lib = _capi.load_library()
    names = declared_functions()
    assert len(names) >= 40
    for n in names:
        assert hasattr(lib, n), "libmogp_b200.so does not export " + n
    assert sorted(_capi.SIGNATURES) == names, "ctypes binding and header disagree"
    assert lib.mogp_version() == 100


def test_no_device_fails_loudly():
    """Without a CUDA device the engine refuses to exist — there is no CPU fallback to fall into."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    from mogp_b200 import _capi
    with pytest.raises(_capi.EngineError):
        _capi.Engine(0)


def test_package_does_not_import_the_oracle():
    import subprocess
    import sys
    code = "import sys; sys.path.insert(0, %r); import mogp_b200, mogp_b200.parallel, mogp_b200.synth; " \
           "assert not any(m == 'oracle' or m.startswith('oracle.') for m in sys.modules), 'oracle imported'" % ROOT
    subprocess.run([sys.executable, "-c", code], check=True)
    for dirpath, _d, files in os.walk(os.path.join(ROOT, "mogp_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".inl")):
                src = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(import|from)\s+oracle\b", src, flags=re.M), f
