"""The C-ABI library loads, exports every symbol include/b200reg.h declares, and the ctypes structures
match the C layout.  No compute calls (no GPU needed)."""
import ctypes as C
import os
import re
import subprocess
import sys
import tempfile

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "b200reg.h")


def _declared_symbols():
    text = open(HEADER).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(b200reg_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol(built):
    from platipy_b200 import _abi

    lib = _abi.load()
    names = _declared_symbols()
    assert len(names) >= 30
    for n in names:
        assert hasattr(lib, n), f"libb200reg.so does not export {n}"
    assert sorted(_abi.SIGNATURES) == names, "ctypes binding and header disagree on the symbol list"
    assert lib.b200reg_abi_version() == 1


def test_ctypes_structs_match_c_layout(built):
    from platipy_b200 import _abi

    src = r'''
    #include <stdio.h>
    #include <stddef.h>
    #include "b200reg.h"
    int main(void) {
      printf("%zu %zu %zu %zu %zu\n", sizeof(b200reg_geom), sizeof(b200reg_transform), sizeof(b200reg_demons_params),
             sizeof(b200reg_demons_stats), sizeof(b200reg_multires_config));
      printf("%zu %zu %zu\n", offsetof(b200reg_transform, d_dvf), offsetof(b200reg_demons_params, max_rms_error),
             offsetof(b200reg_multires_config, demons));
      return 0; }
    '''
    with tempfile.TemporaryDirectory() as d:
        c = os.path.join(d, "t.c")
        open(c, "w").write(src)
        exe = os.path.join(d, "t")
        subprocess.check_call(["/usr/bin/gcc", "-I", os.path.join(ROOT, "include"), c, "-o", exe])
        out = subprocess.check_output([exe]).decode().split()
    sizes = [int(v) for v in out[:5]]
    offs = [int(v) for v in out[5:]]
    assert sizes == [C.sizeof(_abi.Geom), C.sizeof(_abi.Transform), C.sizeof(_abi.DemonsParams), C.sizeof(_abi.DemonsStats),
                     C.sizeof(_abi.MultiresConfig)]
    assert offs == [_abi.Transform.d_dvf.offset, _abi.DemonsParams.max_rms_error.offset, _abi.MultiresConfig.demons.offset]


def test_host_side_entry_points_without_gpu(built):
    """Entry points that are pure host code work without a device; device entry points fail loudly."""
    from platipy_b200 import _abi

    lib = _abi.load()
    g = _abi.make_geom((512, 512, 256), (1.0, 1.0, 1.0), (0, 0, 0), (1, 0, 0, 0, 1, 0, 0, 0, 1))
    out = _abi.Geom()
    assert lib.b200reg_pyramid_geom(C.byref(g), 0, 4.0, C.byref(out)) == 0
    assert tuple(out.size) == (128, 128, 64)
    assert abs(out.spacing[0] - 511.0 / 127.0) < 1e-15 and abs(out.spacing[2] - 255.0 / 63.0) < 1e-15
    # a level that collapses to one voxel is an error (the reference divides by zero there, utils.py:252-255)
    g2 = _abi.make_geom((8, 8, 8), (1.0, 1.0, 1.0), (0, 0, 0), (1, 0, 0, 0, 1, 0, 0, 0, 1))
    assert lib.b200reg_pyramid_geom(C.byref(g2), 0, 8.0, C.byref(out)) == _abi.ERR_ARG
    with pytest.raises(ValueError):
        _abi.check(_abi.ERR_ARG)

    import torch

    if not torch.cuda.is_available():
        ctx = C.c_void_p()
        rc = lib.b200reg_create(0, None, C.byref(ctx))
        assert rc != 0, "creating a context without a GPU must fail"
        with pytest.raises(RuntimeError):
            _abi.check(rc)
        from platipy_b200.engine import Engine

        with pytest.raises(RuntimeError):
            Engine.get()


def test_product_never_imports_the_oracle():
    """The package must not reference oracle/ anywhere (the oracle is a checker, not a fallback)."""
    pkg = os.path.join(ROOT, "platipy_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh")):
                text = open(os.path.join(dirpath, f)).read()
                assert "import oracle" not in text and "from oracle" not in text and "itk_oracle" not in text, f


def test_integration_md_stub_structs_match_the_header():
    """The ctypes stub INTEGRATION.md shows a platipy maintainer (section 3) declares the same structure layouts as include/b200reg.h
    (through platipy_b200._abi, which test_struct_layout_matches_the_c_compiler pins against gcc)."""
    from platipy_b200 import _abi

    text = open(os.path.join(ROOT, "INTEGRATION.md")).read()
    block = re.search(r"```python\nimport ctypes as C, numpy as np\n(.*?)\nlib = C\.CDLL", text, re.S).group(1)
    ns = {"C": C}
    exec(block, ns)
    assert C.sizeof(ns["Geom"]) == C.sizeof(_abi.Geom)
    assert C.sizeof(ns["DemonsParams"]) == C.sizeof(_abi.DemonsParams)
    assert C.sizeof(ns["DemonsStats"]) == C.sizeof(_abi.DemonsStats)
    for name, _ in _abi.DemonsParams._fields_:
        assert getattr(ns["DemonsParams"], name).offset == getattr(_abi.DemonsParams, name).offset, name
