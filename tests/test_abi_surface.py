"""CPU tier: the C-ABI libraries load and export every symbol their headers declare
(no compute calls — there is no GPU here)."""
import ctypes as C
import os
import re

import cases


def _declared(header):
    txt = open(os.path.join(cases.ROOT, "include", header)).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(md[h]?_[a-z_0-9]+)\s*\(", txt)))


def test_libmdgpu_exports_header_symbols(built):
    import subprocess
    subprocess.run(["make", "-s", "-C", os.path.join(cases.ROOT, "methyldackel_b200", "csrc"), "gpu"], check=True, stdout=subprocess.DEVNULL)
    lib = C.CDLL(os.path.join(cases.ROOT, "methyldackel_b200", "lib", "libmdgpu.so"))
    names = _declared("mdgpu.h")
    assert len(names) >= 20
    for n in names:
        assert hasattr(lib, n), n
    lib.md_abi_version.restype = C.c_int
    assert lib.md_abi_version() == 6


def test_libmdhost_exports_header_symbols(built):
    lib = C.CDLL(os.path.join(cases.ROOT, "methyldackel_b200", "lib", "libmdhost.so"))
    for n in [x for x in _declared("mdhost.h") if x.startswith("mdh_")]:
        assert hasattr(lib, n), n


def test_gpu_library_fails_loudly_without_device(built):
    """On a box without a CUDA device md_create must fail with an error, never fall back to the CPU."""
    import torch
    if torch.cuda.is_available():
        return
    from methyldackel_b200 import _abi as A
    g = A.load_gpu()
    cfg = A.default_config()
    assert not g.md_create(C.byref(cfg), 0)
    assert b"CUDA" in g.md_last_error() or b"device" in g.md_last_error()
