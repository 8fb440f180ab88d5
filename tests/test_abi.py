"""The C-ABI library loads on a CPU-only box and exports every symbol include/plip_b200.h declares."""
import ctypes as C
import os
import re

import pytest

from plip_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    hdr = open(os.path.join(ROOT, "include", "plip_b200.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    return sorted(set(re.findall(r"PLIP_API\s+[\w\s\*]+?\b(plip_\w+)\s*\(", hdr)))


def test_header_symbols_all_exported_and_bound():
    names = _declared()
    assert len(names) >= 20
    L = _lib.lib(strict=True)
    for n in names:
        assert hasattr(L, n), f"{n} declared in plip_b200.h but not exported"
        assert n in _lib.SIGNATURES, f"{n} has no ctypes signature"
    assert sorted(_lib.SIGNATURES) == names


def test_abi_version_and_layout_queries():
    L = _lib.lib()
    assert L.plip_abi_version() == 5
    hdr = open(os.path.join(ROOT, "include", "plip_b200.h")).read()
    assert int(re.search(r"#define\s+PLIP_B200_ABI_VERSION\s+(\d+)", hdr).group(1)) == 5

    n = L.plip_weights_num_tensors()
    assert n == 5 + 12 * 10 + 3 + 2 + 12 * 10 + 3
    prev_end = 0
    total_bf16 = 0
    for i in range(n):
        ti = _lib.TensorInfo()
        assert L.plip_weights_tensor_info(i, C.byref(ti)) == 0
        assert ti.offset % 256 == 0 and ti.offset >= prev_end
        assert ti.numel == ti.rows * ti.cols
        prev_end = ti.offset + ti.numel * (2 if ti.dtype == 1 else 4)
        total_bf16 += ti.numel if ti.dtype == 1 else 0
    assert prev_end <= L.plip_weights_blob_bytes() < prev_end + 256
    # all GEMM weights: vision 12*(4*768^2 + 2*768*3072) + patch + proj; text 12*(4*512^2+2*512*2048) + proj
    assert total_bf16 == 12 * (4 * 768 * 768 + 2 * 768 * 3072) + 768 * 3072 + 512 * 768 + \
        12 * (4 * 512 * 512 + 2 * 512 * 2048) + 512 * 512
    assert L.plip_workspace_bytes(1024) > 700e6
    assert L.plip_workspace_bytes(0) == 0


def test_errors_are_codes_not_exceptions():
    L = _lib.lib()
    ti = _lib.TensorInfo()
    assert L.plip_weights_tensor_info(10 ** 6, C.byref(ti)) != 0
    assert "out of range" in _lib.last_error()
    h = C.c_void_p()
    buf = (C.c_char * 16)()
    rc = L.plip_create(C.cast(buf, C.c_void_p), 16, C.c_float(1.0), 0, 8, C.byref(h))
    assert rc != 0 and "blob" in _lib.last_error()
    with pytest.raises(RuntimeError, match="blob"):
        _lib.check(rc, "plip_create")
    assert L.plip_destroy(None) == 0


def test_header_is_plain_c_and_links(tmp_path):
    """include/plip_b200.h is the boundary a C host binds: it must compile as C99 (no C++ / torch types) and the
    shared library must link without CUDA or python on the link line."""
    import shutil
    import subprocess
    gcc = shutil.which("gcc")
    if gcc is None:
        pytest.skip("no gcc")
    src = tmp_path / "host.c"
    src.write_text('#include "plip_b200.h"\n'
                   'int main(void) {\n'
                   '  plip_resize_desc_t d = {0, 1, 1, 224, 224, 0, 0};\n'
                   '  plip_tensor_info_t ti;\n'
                   '  if (sizeof d != 32) return 2;\n'
                   '  if (plip_weights_tensor_info(0, &ti) != 0) return 3;\n'
                   '  if (plip_destroy(0) != 0) return 4;\n'
                   '  return plip_abi_version() == PLIP_B200_ABI_VERSION ? 0 : 1;\n'
                   '}\n')
    exe = tmp_path / "host"
    libdir = os.path.dirname(str(_lib.LIB_PATH))
    subprocess.run([gcc, "-std=c99", "-Wall", "-Wextra", "-pedantic", "-Werror", "-I", os.path.join(ROOT, "include"),
                    str(src), "-L", libdir, "-lplip_b200", f"-Wl,-rpath,{libdir}", "-o", str(exe)], check=True)
    assert subprocess.run([str(exe)]).returncode == 0


def test_argument_validation_needs_no_gpu():
    """Bad arguments are rejected with a code + message before any CUDA call."""
    L = _lib.lib()
    assert L.plip_similarity(None, 1, None, 1, C.c_float(1.0), 0, 0, None, 1, None) != 0
    assert "null argument" in _lib.last_error()
    assert L.plip_l2_normalize(None, 1, 512, None) != 0
    d = _lib.ResizeDesc(0, 10, 10, 224, 224, 0, 0)
    assert L.plip_resize_crop_u8(None, 300, C.byref(d), 1, None, None) != 0
    assert "null argument" in _lib.last_error()
    buf = (C.c_char * 512)()
    addr = C.addressof(buf)
    addr += (-addr) % 4
    assert L.plip_resize_crop_u8(addr, 300, C.byref(d), 0, addr, None) != 0
    assert "positive" in _lib.last_error()
    for bad, msg in [(_lib.ResizeDesc(0, 10, 10, 224, 224, 0, 0), "exceeds"),          # 10x10x3 = 300 > 256
                     (_lib.ResizeDesc(0, 5, 5, 100, 224, 0, 0), "smaller"),
                     (_lib.ResizeDesc(0, 5, 5, 224, 224, 0, 3), "crop origin"),
                     (_lib.ResizeDesc(0, 0, 5, 224, 224, 0, 0), "invalid size")]:
        assert L.plip_resize_crop_u8(addr, 256, C.byref(bad), 1, addr, None) != 0     # rejected on the host
        assert msg in _lib.last_error(), _lib.last_error()
    assert L.plip_resize_crop_u8(addr + 1, 256, C.byref(_lib.ResizeDesc(0, 5, 5, 224, 224, 0, 0)), 1, addr, None) != 0
    assert "aligned" in _lib.last_error()


def test_graft_entry_build_check_follows_the_header():
    """__graft_entry__.build() compares the library's ABI version with the header's (a hard-coded number there broke
    the driver's build check when the ABI moved to 5)."""
    import inspect
    import __graft_entry__ as g
    src = inspect.getsource(g.build)
    assert "PLIP_B200_ABI_VERSION" in src and "== 4" not in src and "== 5" not in src
