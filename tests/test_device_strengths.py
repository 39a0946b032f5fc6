"""Deblocking strengths and motion vectors (rank 3, Baseline streams) derived on the device ( include/b2_recon.h b2_h264_picture.bs_on_device,
h264_bs_kernel).  Through the plugin table the front end leaves the strengths to the device whenever nobody reads the
records on the host, so every golden / reference comparison of tests/test_h264.py already runs on device-derived
strengths; here the two derivations are set against each other on the stream families whose strengths are hardest
(B slices with both lists on one picture, direct modes, 8x8 transform, several slices with per-slice filter control,
intra / I_PCM neighbours), and the extra launch is counted to show that the kernel ran.  The CPU side of the same
claim -- the rules the kernel implements, restated in oracle/recon_h264.c, give the host's strengths on every picture
of every test stream -- is asserted by tests/util.py's picture tap."""
import ctypes as C
import os

import pytest

import streams
import util

NAMES = ["h264_basic", "h264_b_spatial", "h264_high_ip", "h264_cabac_b_mixed", "h264_intra_ipcm", "h264_umv"]


def _decode(data, host_strengths):
    import m2dec_b200
    from m2dec_b200 import _native as L
    md5s, stats = [], (C.c_uint64 * 3)()
    dec = m2dec_b200.M2Decoder("h264", data, host_strengths=host_strengths)
    try:
        dec.run(lambda frm: md5s.append(m2dec_b200.frame_md5(frm)))
        L.lib.b2_context_stats(dec.ctx, C.byref(stats))
    finally:
        dec.close()
    return md5s, stats[2]


@pytest.mark.gpu
@pytest.mark.parametrize("name", NAMES)
def test_device_strengths_equal_host_strengths(name):
    data = open(os.path.join(util.GOLDEN, name + ".264"), "rb").read()
    golden = open(os.path.join(util.GOLDEN, name + ".md5")).read().split()
    dev, dev_launches = _decode(data, False)
    host, host_launches = _decode(data, True)
    assert dev == golden and host == golden
    # h264_bs_kernel per deblocked picture (and h264_mv_kernel + h264_items_kernel per P picture of a Baseline stream) only when the device derives them
    assert host_launches < dev_launches <= host_launches + 3 * len(golden)


@pytest.mark.gpu
@pytest.mark.parametrize("args", [
    "mbw=40 mbh=30 frames=7 seed=31 refs=2 bframes=2 cabac=1 profile=100 transform8x8=1 weighted_bipred=2 intra_pct=10 skip_pct=20 coded_pct=30",
    "mbw=40 mbh=30 frames=7 seed=32 refs=3 bframes=2 profile=77 direct_spatial=0 slices=3 deblock=2 intra_pct=15 direct_pct=40 skip_pct=20",
    "mbw=20 mbh=12 frames=10 seed=34 refs=2 bframes=1 profile=77 direct_8x8_inference=0 cabac=1",
    "mbw=120 mbh=68 frames=3 seed=33 crop_bottom=8 intra_pct=10 skip_pct=15 coded_pct=25",
    "mbw=120 mbh=68 frames=5 seed=35 crop_bottom=8 refs=3 slices=4 deblock=2 mv_range=300 mv_out=60 skip_pct=30",
])
def test_device_strengths_on_writer_streams(args):
    data = streams.generate("h264", args)
    dev, _ = _decode(data, False)
    host, _ = _decode(data, True)
    assert dev == host and dev == util.oracle_md5("h264", data)


def test_oracle_restatements_of_the_device_derivations_run_on_tapped_pictures():
    """CPU side: tests/util.py's picture tap runs oracle_h264_bs_check on every H.264 picture and oracle_h264_mv_check
    on every picture that carries partitions (Baseline streams) -- here it is shown that the latter really happens."""
    data = streams.generate("h264", streams.H264_CASES["h264_refs3"])
    before = util.MV_CHECKED
    md5s = util.oracle_md5("h264", data)
    assert len(md5s) == 12 and util.MV_CHECKED - before == 11      # the IDR picture has no partitions
