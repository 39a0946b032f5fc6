"""Digest mode behind the plugin table (include/b2_recon.h b2_context_set_md5 /
b2_context_frame_md5): the per-frame MD5 of `h264dec -O` (reference src/app/filewrite.h:91-124) computed on the
device must be the one the reference prints -- compared with the committed golden MD5 files, and with the host
hash of the same stream decoded with pixels coming back."""
import ctypes as C
import os

import pytest

import streams
import util

CASES = [("h264", n, "264") for n in ("h264_basic", "h264_umv", "h264_b_spatial", "h264_high_ip", "h264_cabac_b_mixed", "h264_intra_ipcm")] + \
        [("mpeg2", n, "m2v") for n in ("mpeg2_ipb", "mpeg2_crop", "mpeg2_field")]


@pytest.mark.gpu
@pytest.mark.parametrize("codec,name,ext", CASES)
def test_device_md5_matches_golden(codec, name, ext):
    import m2dec_b200
    data = open(os.path.join(util.GOLDEN, name + "." + ext), "rb").read()
    golden = open(os.path.join(util.GOLDEN, name + ".md5")).read().split()
    assert m2dec_b200.decode_file(codec, data, device_md5=True) == golden


@pytest.mark.gpu
def test_device_md5_cropped_1080p_and_nothing_but_digests_comes_back():
    """The bench geometry (1920x1088 coded, 1080 displayed): digests equal the host hash of the frames of a normal
    decode; in digest mode 16 bytes per frame cross PCIe and the frame buffers stay untouched."""
    import m2dec_b200
    from m2dec_b200 import _native as L
    data = streams.generate("h264", "mbw=120 mbh=68 frames=4 seed=7 crop_bottom=8 intra_pct=10 skip_pct=15 coded_pct=25")
    want = m2dec_b200.decode_file("h264", data)
    got, stats = [], (C.c_uint64 * 3)()
    dec = m2dec_b200.M2Decoder("h264", data, device_md5=True)
    try:
        dec.run(lambda frm: got.append(dec.frame_digest(frm)))
        L.lib.b2_context_stats(dec.ctx, C.byref(stats))
    finally:
        dec.close()
    assert got == want and len(got) == 4
    assert stats[1] == 16 * 4, list(stats)


def test_digest_entry_points_refuse_without_digest_mode():
    """No device, digest mode never switched on: the call fails instead of returning something."""
    import m2dec_b200
    from m2dec_b200 import _native as L
    data = open(os.path.join(util.GOLDEN, "h264_basic.264"), "rb").read()
    dec = m2dec_b200.M2Decoder("h264", data, tap=util.oracle_tap, device_off=True)
    try:
        frames = []
        dec.run(lambda frm: frames.append(L.FrameT.from_buffer_copy(frm)))
        out = (C.c_uint8 * 16)()
        assert frames and L.lib.b2_context_frame_md5(dec.ctx, C.byref(frames[0]), out) < 0
    finally:
        dec.close()
