"""Kernel replay (b2_replay_*, the path bench.py's `value` is measured on): several streams of one
geometry but very different tools are reconstructed together, picture #t of every lane per batched
launch, and every picture must be bit-exact against the CPU oracle (and through it the reference)."""
import hashlib

import numpy as np
import pytest

import streams
import util

GEOM = "mbw=10 mbh=8 frames=10"
LANES = [
    "seed=201",
    "seed=202 refs=3 slices=3 deblock=2",
    "seed=203 intra_pct=50 ipcm_pct=10",
    "seed=204 refs=2 bframes=2 profile=77 weighted_bipred=1 weighted_pred=1",
    "seed=205 cabac=1 profile=100 transform8x8=1 refs=2 bframes=1",
    "seed=206 qp=40 alpha_div2=3 beta_div2=4 chroma_qp_offset=4",
    "seed=207 qp=30 mv_range=300 mv_out=60 skip_pct=40",
    "seed=208 deblock=0",                                       # loop filter off, paired with a picture that has it on
    "seed=209 qp=20 coded_pct=60 alpha_div2=-2 beta_div2=1",    # an odd number of lanes: the last one has no partner picture
]


def _md5(y, c):
    h = hashlib.md5()
    h.update(np.ascontiguousarray(y).tobytes())
    h.update(np.ascontiguousarray(c).tobytes())
    return h.hexdigest()


@pytest.mark.gpu
@pytest.mark.parametrize("nlanes", [len(LANES), 8, 2, 1])
def test_replay_matches_oracle(nlanes):
    import m2dec_b200
    data = [streams.generate("h264", GEOM + " " + a) for a in LANES[:nlanes]]
    want = [sorted(util.oracle_md5("h264", d)) for d in data]
    rp = m2dec_b200.Replay("h264", 160, 128, nlanes, pool_frames=24)
    try:
        for lane, d in enumerate(data):
            assert rp.load_stream(lane, d) == 10
        for _ in range(2):          # a second run over the same resident pictures must give the same frames
            rp.run()
            rp.sync()
            for lane in range(nlanes):
                got = sorted(_md5(*rp.download(lane, idx)) for idx in rp.outputs[lane])
                assert got == want[lane], "lane %d (%s)" % (lane, LANES[lane])
    finally:
        rp.close()


@pytest.mark.gpu
def test_pageable_and_page_locked_frame_buffers_agree():
    """The plugin path copies a picture out as soon as it is queued when the application's frame
    buffers are page-locked (b2_host_alloc), and on demand when they are pageable: same frames."""
    import m2dec_b200
    from m2dec_b200 import decoder
    data = streams.generate("h264", GEOM + " " + LANES[3])
    want = util.oracle_md5("h264", data)
    assert m2dec_b200.decode_file("h264", data) == want          # page-locked (default when a device is present)
    saved, decoder._frames_pool = decoder._frames_pool, {}
    orig = decoder.Frames.__init__
    try:
        decoder.Frames.__init__ = lambda self, w, h, n, s, pinned=True, cuda_device=-1: orig(self, w, h, n, s, False, cuda_device)
        assert m2dec_b200.decode_file("h264", data) == want      # pageable numpy buffers
    finally:
        decoder.Frames.__init__ = orig
        decoder._frames_pool = saved


@pytest.mark.gpu
@pytest.mark.parametrize("geom,window", [("mbw=8 mbh=6 frames=5", (0, 0, 128, 96)), ("mbw=8 mbh=6 frames=5", (0, 2, 128, 90)),
                                         ("mbw=10 mbh=8 frames=5", (0, 0, 160, 128)), ("mbw=10 mbh=8 frames=5", (6, 4, 150, 120))])
def test_device_md5_matches_hashlib(geom, window):
    """b2_replay_md5 (one thread per picture) against hashlib over the downloaded, cropped planes:
    the word-aligned path (width a multiple of 64) and the byte path, with and without cropping."""
    import m2dec_b200
    left, top, w, h = window
    mbw, mbh = int(geom.split()[0][4:]), int(geom.split()[1][4:])
    rp = m2dec_b200.Replay("h264", mbw * 16, mbh * 16, 3, pool_frames=24)
    try:
        for lane in range(3):
            assert rp.load_stream(lane, streams.generate("h264", geom + " seed=%d" % (300 + lane))) == 5
        rp.run()
        rp.sync()
        pics = [(lane, idx) for lane in range(3) for idx in rp.outputs[lane]]
        want = []
        for lane, idx in pics:
            y, c = rp.download(lane, idx)
            want.append(_md5(y[top:top + h, left:left + w], c[top // 2:top // 2 + h // 2, left:left + w]))
        assert rp.md5(pics, width=w, height=h, left=left, top=top) == want
    finally:
        rp.close()
