"""The C ABI (include/*.h): every declared entry point is exported by libm2dec_b200.so, the plugin
tables have the reference's layout, and without a CUDA device the product path refuses to decode
instead of falling back to anything (there is no CPU reconstruction in the library)."""
import ctypes as C
import os
import re

import pytest

import streams
import util

INCLUDE = os.path.join(util.ROOT, "include")


def _declared():
    fns, data = [], []
    for name in sorted(os.listdir(INCLUDE)):
        text = open(os.path.join(INCLUDE, name)).read()
        text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
        for line in text.splitlines():
            line = line.strip()
            if line.startswith("typedef") or not line.endswith(";"):
                continue
            m = re.match(r"extern\s+const\s+\w+\s*\*\s*const\s+(\w+)\s*;", line)
            if m:
                data.append(m.group(1))
                continue
            m = re.match(r"(?:const\s+)?\w[\w\s]*?[\s\*]+(\w+)\s*\(.*\)\s*;", line)
            if m and not line.startswith(("return", "extern \"C\"")) and "(*" not in line.split("(")[0]:
                fns.append(m.group(1))
    return fns, data


def test_every_declared_symbol_is_exported():
    import m2dec_b200
    fns, data = _declared()
    assert len(fns) >= 35 and set(data) == {"m2d_gpu_func", "h264d_gpu_func"}, (fns, data)
    missing = [f for f in fns + data if not hasattr(m2dec_b200.lib, f)]
    assert not missing, missing


def test_plugin_tables_have_the_reference_layout():
    """m2d_func_table_t: context_size + seven entry points, all set (reference src/lib/m2d.h:66-75)."""
    from m2dec_b200 import _native as L
    for name in ("m2d_gpu_func", "h264d_gpu_func"):
        t = L.func_table(name)
        assert t.context_size >= 32
        for field, _ in L.FuncTable._fields_[1:]:
            assert C.cast(getattr(t, field), C.c_void_p).value, (name, field)
    assert C.sizeof(L.FrameT) == 40 and C.sizeof(L.InfoT) == 24      # m2d_frame_t / m2d_info_t on LP64


def test_no_cuda_device_means_no_decode():
    """Without a device: no page-locked memory, no session, and the decoder reports -4 rather than
    producing frames some other way.  (Skipped where a GPU is present: the GPU tests cover that side.)"""
    import m2dec_b200
    from m2dec_b200 import _native as L
    if L.lib.b2_device_count() > 0:
        pytest.skip("a CUDA device is present")
    assert not L.lib.b2_host_alloc(1 << 20, -1)
    sess = C.c_void_p()
    L.lib.b2_session_create.argtypes = [C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_void_p)]
    assert L.lib.b2_session_create(0, 2, 160, 128, 4, C.byref(sess)) < 0 and not sess.value
    assert L.lib.b2_last_error()
    data = streams.generate("h264", streams.H264_CASES["h264_basic"])
    with pytest.raises(RuntimeError):
        m2dec_b200.decode_file("h264", data)
