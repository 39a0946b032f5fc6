"""MPEG-2 path: product parse + CPU oracle vs the reference's MD5 lines (CPU tests), and
product parse + CUDA reconstruction through the plugin table vs the same (GPU tests)."""
import os

import pytest

import util

CASES = ["cv2_mpeg2_352x288"]


def _load(name):
    data = open(os.path.join(util.GOLDEN, name + ".m2v"), "rb").read()
    md5 = open(os.path.join(util.GOLDEN, name + ".md5")).read().split()
    return data, md5


@pytest.mark.parametrize("name", CASES)
def test_oracle_matches_reference_golden(name):
    data, golden = _load(name)
    assert util.oracle_md5("mpeg2", data) == golden


@pytest.mark.parametrize("name", CASES)
def test_golden_is_what_the_reference_says(name):
    if not util.have_ref():
        pytest.skip("oracle/_ref not built")
    data, golden = _load(name)
    assert util.ref_md5(data, "m2v") == golden


@pytest.mark.gpu
@pytest.mark.parametrize("name", CASES)
def test_cuda_matches_reference_golden(name):
    import m2dec_b200
    data, golden = _load(name)
    assert m2dec_b200.decode_file("mpeg2", data) == golden


import streams  # noqa: E402


def _golden(name):
    data = open(os.path.join(util.GOLDEN, name + ".m2v"), "rb").read()
    md5 = open(os.path.join(util.GOLDEN, name + ".md5")).read().split()
    return data, md5


@pytest.mark.parametrize("name", streams.MPEG2_GOLDEN)
def test_writer_is_deterministic(name):
    data, _ = _golden(name)
    assert streams.generate("mpeg2", streams.MPEG2_CASES[name]) == data


@pytest.mark.parametrize("name", streams.MPEG2_GOLDEN)
def test_oracle_matches_golden(name):
    data, golden = _golden(name)
    assert util.oracle_md5("mpeg2", data) == golden


@pytest.mark.parametrize("name", sorted(streams.MPEG2_CASES))
def test_oracle_matches_compiled_reference(name):
    if not util.have_ref():
        pytest.skip("oracle/_ref not built")
    data = streams.generate("mpeg2", streams.MPEG2_CASES[name])
    ref = util.ref_md5(data, "m2v")
    assert len(ref) > 0
    assert util.oracle_md5("mpeg2", data) == ref


@pytest.mark.gpu
@pytest.mark.parametrize("name", streams.MPEG2_GOLDEN)
def test_cuda_matches_golden_writer_streams(name):
    import m2dec_b200
    data, golden = _golden(name)
    assert m2dec_b200.decode_file("mpeg2", data) == golden


@pytest.mark.gpu
@pytest.mark.parametrize("name", sorted(streams.MPEG2_CASES))
def test_cuda_matches_oracle_and_reference(name):
    import m2dec_b200
    data = streams.generate("mpeg2", streams.MPEG2_CASES[name])
    gpu = m2dec_b200.decode_file("mpeg2", data)
    assert gpu == util.oracle_md5("mpeg2", data)
    if util.have_ref():
        assert gpu == util.ref_md5(data, "m2v")


@pytest.mark.gpu
def test_cuda_1080p_mp_hl():
    """BASELINE config 4: MPEG-2 MP@HL 1920x1080 I/P/B, checked against the compiled reference
    (or the oracle where the reference binary is absent)."""
    import m2dec_b200
    data = streams.generate("mpeg2", "mbw=120 mbh=68 width=1920 height=1080 frames=7 seed=40")
    gpu = m2dec_b200.decode_file("mpeg2", data)
    want = util.ref_md5(data, "m2v") if util.have_ref() else util.oracle_md5("mpeg2", data)
    assert len(gpu) == 7 and gpu == want


# ---- frame-out: the held frame (four buffers or more) and the reference's own timing (three buffers) ----------------
@pytest.mark.parametrize("name", streams.MPEG2_GOLDEN)
@pytest.mark.parametrize("outbuf", [-1, 0, 3])
def test_frame_out_order_does_not_depend_on_the_pool(name, outbuf):
    """get_info asks for four frame buffers so that a due frame can wait one picture for the device; an application
    that hands over three (the reference's number) gets every frame at once, one that hands over more gets the same
    frames in the same order."""
    data, golden = _golden(name)
    assert util.oracle_md5("mpeg2", data, outbuf=outbuf) == golden


@pytest.mark.gpu
@pytest.mark.parametrize("name", ["mpeg2_ipb", "mpeg2_field"])
@pytest.mark.parametrize("outbuf", [-1, 0, 3])
def test_cuda_frame_out_order_does_not_depend_on_the_pool(name, outbuf):
    import m2dec_b200
    data, golden = _golden(name)
    assert m2dec_b200.decode_file("mpeg2", data, outbuf=outbuf) == golden
    assert m2dec_b200.decode_file("mpeg2", data, outbuf=outbuf, device_md5=True) == golden


@pytest.mark.parametrize("name", streams.MPEG2_GOLDEN)
@pytest.mark.parametrize("outbuf", [-1, 0])
def test_frame_out_order_in_emptify_mode(name, outbuf):
    """The reference application's other loop (m2decoder.h:149-152: collect every frame that is due before decoding on)
    sees the same frames in the same order."""
    from m2dec_b200.decoder import M2Decoder, frame_md5
    data, golden = _golden(name)
    md5s = []
    dec = M2Decoder("mpeg2", data, tap=util.oracle_tap, device_off=True, outbuf=outbuf)
    try:
        dec.run(lambda f: md5s.append(frame_md5(f)), emptify=True)
    finally:
        dec.close()
    assert md5s == golden
