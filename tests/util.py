"""Shared test helpers: the compiled reference (oracle/_ref), the CPU oracle (oracle/_build)
and the plumbing that feeds the same SoA pictures to the oracle through the picture tap."""
import ctypes as C
import os
import subprocess
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF_BIN = os.path.join(ROOT, "oracle", "_ref", "h264dec")
ORACLE_SO = os.path.join(ROOT, "oracle", "_build", "liboracle.so")
GOLDEN = os.path.join(ROOT, "tests", "golden")

_oracle = None


def have_ref():
    return os.path.exists(REF_BIN)


def oracle():
    global _oracle
    if _oracle is None:
        if not os.path.exists(ORACLE_SO):
            subprocess.check_call(["make", "-C", os.path.join(ROOT, "oracle"), "oracle"],
                                  stdout=subprocess.DEVNULL)
        _oracle = C.CDLL(ORACLE_SO)
    return _oracle


def ref_md5(data, ext, args=()):
    """MD5 lines written by the unmodified reference CLI (`h264dec -O`).  MPEG-2 streams get
    32 bytes of 0xFF appended (works around the reference's EOF longjmp)."""
    with tempfile.TemporaryDirectory() as d:
        path = os.path.join(d, "s." + ext)
        with open(path, "wb") as f:
            f.write(data + (b"\xff" * 32 if ext == "m2v" else b""))
        subprocess.run([REF_BIN, "-O", *args, path], cwd=d, stdout=subprocess.DEVNULL,
                       stderr=subprocess.DEVNULL, timeout=600)
        out = os.path.join(d, "s.out")
        if not os.path.exists(out):
            return []
        with open(out, "rb") as f:
            return [l.strip().decode() for l in f.read().split(b"\r\n") if l.strip()]


MV_CHECKED = 0          # pictures whose partitions were checked against their motion records (oracle_h264_mv_check)


def oracle_tap(dec, codec, pic):
    """Picture tap that reconstructs on the CPU oracle into the decoder's host frames."""
    fr = dec.frames
    n = fr.num
    luma = (C.c_void_p * n)(*[fr.array[i].luma for i in range(n)])
    chroma = (C.c_void_p * n)(*[fr.array[i].chroma for i in range(n)])
    o = oracle()
    # every picture the front end emits must pass the structural check the device entry points run
    from m2dec_b200 import _native as L
    check = L.lib.b2_mpeg2_picture_check if codec == 1 else L.lib.b2_h264_picture_check
    assert check(C.c_void_p(pic), n) == 0, (L.lib.b2_last_error() or b"").decode()
    if codec == 1:
        return o.oracle_mpeg2_picture(C.c_void_p(pic), luma, chroma)
    # the deblocking strengths the device would derive from the same records (oracle restatement of h264_bs_kernel's
    # rules) must be the ones the host front end filled in: every tapped picture of every test stream checks it
    bad = o.oracle_h264_bs_check(C.c_void_p(pic))
    assert bad == 0, "%d macroblocks: strengths derived from the records differ from the host's" % bad
    # likewise the motion vectors the device would predict from the partitions (Baseline streams carry them next to the
    # records when a tap looks on): -1 = no partitions in this picture
    bad = o.oracle_h264_mv_check(C.c_void_p(pic))
    assert bad <= 0, "%d motion records differ from the vectors predicted from the partitions" % bad
    global MV_CHECKED
    MV_CHECKED += bad == 0
    return o.oracle_h264_picture(C.c_void_p(pic), luma, chroma)


def oracle_md5(codec, data, **kw):
    """Product parse -> SoA -> CPU oracle reconstruction -> per-frame MD5 (no GPU involved)."""
    import m2dec_b200
    return m2dec_b200.decode_file(codec, data, tap=oracle_tap, device_off=True, **kw)
