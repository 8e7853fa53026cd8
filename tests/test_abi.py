"""The C ABI: include/lux_b200.h, the numpy layout and the built library agree (no GPU needed)."""
import ctypes
import os
import re
import subprocess
import tempfile

import numpy as np
import pytest

from luxpythonenvgym_b200 import _lib, build
from luxpythonenvgym_b200 import state as S

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "lux_b200.h")


def test_header_declares_exactly_the_exported_symbols():
    text = open(HEADER).read()
    declared = set(re.findall(r"\b(lux_b200_\w+)\s*\(", text))
    assert declared == set(_lib.EXPORTS)


def test_library_builds_loads_and_exports_everything():
    build.build()
    lib = _lib.load()
    for name in _lib.EXPORTS:
        assert hasattr(lib, name), name
    assert lib.lux_b200_version() >= 100
    assert lib.lux_b200_step_smem_bytes(32, 252, 255, 384, 256) < 64 * 1024
    assert lib.lux_b200_set_option(b"no_such_option", 1) == _lib.LUX_E_ARG


def test_struct_layout_matches_numpy_dtypes():
    fields = {"LuxHeader": S.header_dtype, "LuxCell": S.cell_dtype, "LuxUnit": S.unit_dtype, "LuxCity": S.city_dtype}
    lines = ['#include <stdio.h>', '#include <stddef.h>', '#include "%s"' % HEADER, "int main(void){"]
    for struct, dt in fields.items():
        lines.append('printf("%s.sizeof %%zu\\n", sizeof(%s));' % (struct, struct))
        for name in dt.names:
            lines.append('printf("%s.%s %%zu\\n", offsetof(%s, %s));' % (struct, name, struct, name))
    lines.append("return 0;}")
    with tempfile.TemporaryDirectory() as d:
        src, exe = os.path.join(d, "l.c"), os.path.join(d, "l")
        open(src, "w").write("\n".join(lines))
        subprocess.run(["gcc", "-o", exe, src], check=True)
        out = subprocess.run([exe], capture_output=True, text=True, check=True).stdout
    got = dict(line.split() for line in out.strip().splitlines())
    for struct, dt in fields.items():
        assert int(got[struct + ".sizeof"]) == dt.itemsize
        for name in dt.names:
            assert int(got["%s.%s" % (struct, name)]) == dt.fields[name][1], (struct, name)


def test_no_device_is_an_error_not_a_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a device is present")
    lib = _lib.load()
    assert lib.lux_b200_device_count() == 0
    n, size, caps = 2, 12, S.default_caps(12)
    bufs = [np.zeros(n * w + 16, dtype=np.uint8) for w in (128, 144 * 8, caps["units"] * 16, caps["cities"] * 16,
                                                           caps["tiles"] * 2, caps["res"] * 4)]
    ptrs = [(b.ctypes.data + 15) & ~15 for b in bufs]
    desc = _lib.LuxBatch(n, size, caps["units"], caps["cities"], caps["tiles"], caps["res"], *ptrs)
    ua = np.zeros(n * caps["units"] + 4, dtype=np.uint32)
    ta = np.zeros(n * caps["tiles"] + 16, dtype=np.uint8)
    rc = lib.lux_b200_step(ctypes.byref(desc), (ua.ctypes.data + 15) & ~15, (ta.ctypes.data + 15) & ~15, 0, None)
    assert rc == _lib.LUX_E_NO_DEVICE
    with pytest.raises(RuntimeError):
        from luxpythonenvgym_b200.batched import BatchedGame
        BatchedGame(2, 12)
