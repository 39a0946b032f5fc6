"""H.264 path.  CPU tests: product parse + CPU oracle vs the reference (committed golden MD5s,
and the compiled reference itself where oracle/_ref exists).  GPU tests: product parse + CUDA
reconstruction through the plugin table vs the same."""
import os

import pytest

import streams
import util


def _golden(name):
    data = open(os.path.join(util.GOLDEN, name + ".264"), "rb").read()
    md5 = open(os.path.join(util.GOLDEN, name + ".md5")).read().split()
    return data, md5


@pytest.mark.parametrize("name", streams.H264_GOLDEN)
def test_writer_is_deterministic(name):
    data, _ = _golden(name)
    assert streams.generate("h264", streams.H264_CASES[name]) == data


@pytest.mark.parametrize("name", streams.H264_GOLDEN)
def test_oracle_matches_golden(name):
    data, golden = _golden(name)
    assert util.oracle_md5("h264", data) == golden


@pytest.mark.parametrize("name", sorted(streams.H264_CASES))
def test_oracle_matches_compiled_reference(name):
    if not util.have_ref():
        pytest.skip("oracle/_ref not built")
    data = streams.generate("h264", streams.H264_CASES[name])
    ref = util.ref_md5(data, "264")
    assert len(ref) > 0
    assert util.oracle_md5("h264", data) == ref


@pytest.mark.gpu
@pytest.mark.parametrize("name", streams.H264_GOLDEN)
def test_cuda_matches_golden(name):
    import m2dec_b200
    data, golden = _golden(name)
    assert m2dec_b200.decode_file("h264", data) == golden


@pytest.mark.gpu
@pytest.mark.parametrize("name", sorted(streams.H264_CASES))
def test_cuda_matches_oracle_and_reference(name):
    import m2dec_b200
    data = streams.generate("h264", streams.H264_CASES[name])
    gpu = m2dec_b200.decode_file("h264", data)
    assert gpu == util.oracle_md5("h264", data)
    if util.have_ref():
        assert gpu == util.ref_md5(data, "264")


def _ffmpeg_luma(data, w, h, n):
    """Luma planes from OpenCV's bundled ffmpeg H.264 decoder (an independent implementation)."""
    import tempfile
    import cv2
    import numpy as np
    with tempfile.NamedTemporaryFile(suffix=".264") as f:
        f.write(data)
        f.flush()
        cap = cv2.VideoCapture(f.name)
        cap.set(cv2.CAP_PROP_CONVERT_RGB, 0)
        out = []
        for _ in range(n):
            ok, frm = cap.read()
            if not ok:
                break
            out.append(np.asarray(frm).reshape(-1, w)[:h].copy())
    return out


@pytest.mark.parametrize("name", sorted(streams.H264_REFERENCE_DEVIATES))
def test_degenerate_sizes_match_ffmpeg(name):
    cv2 = pytest.importorskip("cv2")  # noqa: F841
    import m2dec_b200
    args = streams.H264_REFERENCE_DEVIATES[name]
    kv = dict(a.split("=") for a in args.split())
    w, h = int(kv["mbw"]) * 16, int(kv["mbh"]) * 16
    data = streams.generate("h264", args)
    ours = []
    m2dec_b200.decode_file("h264", data, post=lambda f: ours.append(m2dec_b200.crop_frame(f)[0].copy()),
                           tap=util.oracle_tap, device_off=True)
    theirs = _ffmpeg_luma(data, w, h, len(ours))
    assert len(theirs) == len(ours) > 0
    for a, b in zip(ours, theirs):
        assert (a == b).all()


@pytest.mark.gpu
@pytest.mark.parametrize("name", sorted(streams.H264_REFERENCE_DEVIATES))
def test_cuda_matches_oracle_degenerate(name):
    import m2dec_b200
    data = streams.generate("h264", streams.H264_REFERENCE_DEVIATES[name])
    assert m2dec_b200.decode_file("h264", data) == util.oracle_md5("h264", data)


@pytest.mark.parametrize("name", sorted(streams.H264_ORACLE_ONLY))
def test_oracle_only_streams_decode(name):
    """Streams the reference cannot parse (see streams.H264_ORACLE_ONLY): the product parse must
    still deliver every picture."""
    args = streams.H264_ORACLE_ONLY[name]
    frames = int(dict(a.split("=") for a in args.split())["frames"])
    assert len(util.oracle_md5("h264", streams.generate("h264", args))) == frames


@pytest.mark.gpu
@pytest.mark.parametrize("name", sorted(streams.H264_ORACLE_ONLY))
def test_cuda_matches_oracle_only(name):
    import m2dec_b200
    data = streams.generate("h264", streams.H264_ORACLE_ONLY[name])
    assert m2dec_b200.decode_file("h264", data) == util.oracle_md5("h264", data)


@pytest.mark.gpu
def test_cuda_concurrent_streams():
    """Many decoder contexts on host threads at once (one CUDA stream each): every stream must
    still be bit-exact."""
    import threading
    import m2dec_b200
    names = sorted(streams.H264_CASES)
    data = {n: streams.generate("h264", streams.H264_CASES[n]) for n in names}
    want = {n: util.oracle_md5("h264", data[n]) for n in names}
    got = {}

    def work(n):
        got[n] = m2dec_b200.decode_file("h264", data[n])

    for _ in range(2):
        ths = [threading.Thread(target=work, args=(n,)) for n in names]
        for t in ths:
            t.start()
        for t in ths:
            t.join()
        for n in names:
            assert got[n] == want[n], n


@pytest.mark.gpu
def test_cuda_1080p_baseline():
    """BASELINE config 2: H.264 Baseline CAVLC 1920x1080 (coded 1920x1088, crop 8) I/P."""
    import m2dec_b200
    data = streams.generate("h264", "mbw=120 mbh=68 frames=6 seed=41 crop_bottom=8 refs=2")
    gpu = m2dec_b200.decode_file("h264", data)
    want = util.ref_md5(data, "264") if util.have_ref() else util.oracle_md5("h264", data)
    assert len(gpu) == 6 and gpu == want


@pytest.mark.gpu
def test_cuda_1080p_high_cabac_b():
    """BASELINE config 3: H.264 CABAC 1920x1080 with B pictures, weighted prediction, both direct
    modes' machinery and the High-profile 8x8 transform / Intra8x8."""
    import m2dec_b200
    data = streams.generate("h264", "mbw=120 mbh=68 frames=7 seed=43 crop_bottom=8 refs=2 bframes=2 cabac=1 profile=100 "
                                    "transform8x8=1 weighted_pred=1 weighted_bipred=2 intra_pct=10")
    gpu = m2dec_b200.decode_file("h264", data)
    want = util.ref_md5(data, "264") if util.have_ref() else util.oracle_md5("h264", data)
    assert len(gpu) == 7 and gpu == want


def test_oracle_1080p_high_cabac_b_slice():
    """Same configuration, 1920x1088 coded, fewer frames: product parse + CPU oracle vs the compiled reference."""
    if not util.have_ref():
        pytest.skip("oracle/_ref not built")
    data = streams.generate("h264", "mbw=120 mbh=68 frames=4 seed=44 crop_bottom=8 refs=2 bframes=2 cabac=1 profile=100 "
                                    "transform8x8=1 weighted_pred=1 weighted_bipred=1 direct_spatial=0")
    assert util.oracle_md5("h264", data) == util.ref_md5(data, "264")


@pytest.mark.gpu
def test_cuda_parked_sessions_start_clean():
    """A stream's device session is parked when the stream ends and handed to the next stream of the same geometry
    (session.cu: session_park): streams of one geometry decoded back to back -- frames, then digests, then frames again,
    with another geometry in between -- give the MD5s a first decode gives, and b2_cache_trim frees the parked sessions."""
    import m2dec_b200
    from m2dec_b200 import _native as L
    names = ["h264_basic", "h264_refs3", "h264_b_spatial", "h264_cabac_ip"]
    want = {n: _golden(n)[1] for n in names}
    for rnd in range(3):
        for n in names:
            data = _golden(n)[0]
            assert m2dec_b200.decode_file("h264", data, device_md5=(rnd == 1)) == want[n], (rnd, n)
    L.lib.b2_cache_trim()
    for n in names:
        assert m2dec_b200.decode_file("h264", _golden(n)[0]) == want[n], n
