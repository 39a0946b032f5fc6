"""Damaged input must end in an error return, never in a crash, and whatever pictures the front end still
emits must pass the structural check the device entry points run (b2_*_picture_check): the kernels index
memory with what the records say.  Product parse only (no GPU, no oracle: the pictures are garbage)."""
import ctypes as C
import os
import random

import pytest

import m2dec_b200
import streams
import util
from m2dec_b200 import _native as L
from m2dec_b200.decoder import M2Decoder


def _decode(codec, data, **kw):
    """Parse-only decode; returns (frames out, pictures emitted, pictures that failed the check)."""
    stats = {"pics": 0, "bad": 0}

    def tap(dec, c, pic):
        stats["pics"] += 1
        check = L.lib.b2_mpeg2_picture_check if c == 1 else L.lib.b2_h264_picture_check
        if check(C.c_void_p(pic), dec.frames.num) < 0:
            stats["bad"] += 1
        return 0

    out = [0]
    dec = M2Decoder(codec, data, tap=tap, device_off=True, **kw)
    try:
        dec.run(lambda f: out.__setitem__(0, out[0] + 1))
    except RuntimeError:
        pass                    # the application gives up: set_frames refused the geometry of a damaged header
    finally:
        dec.close()
    return out[0], stats["pics"], stats["bad"]


def _golden(name, ext):
    return open(os.path.join(util.GOLDEN, name + ext), "rb").read()


H264 = ["h264_basic", "h264_b3_explicit_slices", "h264_cabac_b_mixed", "h264_high_ip", "h264_intra_ipcm"]
MPEG2 = ["mpeg2_ipb", "mpeg2_field", "cv2_mpeg2_352x288"]


@pytest.mark.parametrize("name", H264)
def test_h264_truncated(name):
    data = _golden(name, ".264")
    full = _decode("h264", data)
    rng = random.Random(1)
    for cut in sorted(rng.sample(range(1, len(data)), 40)):
        frames, pics, bad = _decode("h264", data[:cut])
        assert frames <= full[0] and bad == 0, cut


@pytest.mark.parametrize("name", H264)
def test_h264_corrupted(name):
    data = _golden(name, ".264")
    rng = random.Random(2)
    for trial in range(60):
        d = bytearray(data)
        for _ in range(rng.choice([1, 1, 2, 8, 64])):
            pos = rng.randrange(len(d))
            d[pos] = rng.randrange(256) if rng.random() < 0.5 else d[pos] ^ (1 << rng.randrange(8))
        frames, pics, bad = _decode("h264", bytes(d), chunk=rng.choice([0, 0, 7, 1000]))
        assert bad == 0, trial


@pytest.mark.parametrize("name", MPEG2)
def test_mpeg2_truncated_and_corrupted(name):
    data = _golden(name, ".m2v")
    rng = random.Random(3)
    for cut in sorted(rng.sample(range(1, len(data)), 30)):
        assert _decode("mpeg2", data[:cut])[2] == 0, cut
    for trial in range(60):
        d = bytearray(data)
        for _ in range(rng.choice([1, 1, 2, 8, 64])):
            pos = rng.randrange(len(d))
            d[pos] = rng.randrange(256) if rng.random() < 0.5 else d[pos] ^ (1 << rng.randrange(8))
        assert _decode("mpeg2", bytes(d), chunk=rng.choice([0, 0, 7, 1000]))[2] == 0, trial


def test_mpeg2_truncated_last_picture_is_flushed():
    """End of data inside the last picture: it is still submitted (frame-out must not hand out stale pixels)."""
    data = _golden("mpeg2_ipb", ".m2v")
    full = _decode("mpeg2", data)
    last = data.rfind(b"\x00\x00\x01\x00")          # last picture_start_code
    cut = last + (len(data) - last) * 2 // 3         # somewhere inside its slices
    frames, pics, bad = _decode("mpeg2", data[:cut])
    assert pics == full[1] and bad == 0


# ---- hand-made H.264 headers ------------------------------------------------------------------------
class Bits:
    def __init__(self):
        self.b = []

    def u(self, n, v):
        for i in range(n - 1, -1, -1):
            self.b.append((v >> i) & 1)
        return self

    def ue(self, v):
        v += 1
        n = v.bit_length()
        return self.u(n - 1, 0).u(n, v)

    def se(self, v):
        return self.ue(2 * v - 1 if v > 0 else -2 * v)

    def nal(self, hdr):
        bits = self.b + [1]
        bits += [0] * (-len(bits) % 8)
        body = bytes(int("".join(map(str, bits[i:i + 8])), 2) for i in range(0, len(bits), 8))
        out = bytearray()
        zeros = 0
        for c in body:                     # emulation prevention
            if zeros >= 2 and c <= 3:
                out.append(3)
                zeros = 0
            out.append(c)
            zeros = zeros + 1 if c == 0 else 0
        return b"\x00\x00\x01" + bytes([hdr]) + bytes(out)


def sps(w_mbs=4, h_mbs=3, log2_fn=0, poc_type=0, log2_poc=0, refs=1, profile=66, sps_id=0):
    b = Bits().u(8, profile).u(8, 0).u(8, 30).ue(sps_id).ue(log2_fn).ue(poc_type)
    if poc_type == 0:
        b.ue(log2_poc)
    b.ue(refs).u(1, 0).ue(w_mbs - 1).ue(h_mbs - 1).u(1, 1).u(1, 1).u(1, 0).u(1, 0)
    return b.nal(0x67)


def pps(pps_id=0, sps_id=0, init_qp=0):
    return Bits().ue(pps_id).ue(sps_id).u(1, 0).u(1, 0).ue(0).ue(0).ue(0).u(1, 0).u(2, 0).se(init_qp).se(0).se(0).u(1, 1).u(1, 0).u(1, 0).nal(0x68)


def idr_slice(first_mb, n_mbs, qp_delta=0):
    b = Bits().ue(first_mb).ue(7).ue(0).u(4, 0).ue(0).u(4, 0).u(1, 0).u(1, 0).se(qp_delta).ue(1)
    for _ in range(n_mbs):                 # I_PCM macroblocks: mb_type 25, alignment, 384 bytes
        b.ue(25)
        b.b += [0] * (-len(b.b) % 8)
        for _ in range(384):
            b.u(8, 0x55)
    return b.nal(0x65)


def test_handmade_stream_decodes():
    data = sps() + pps() + idr_slice(0, 12)
    assert _decode("h264", data) == (1, 1, 0)


@pytest.mark.parametrize("first_mb", [12, 13, 1000, 2 ** 31 - 1, 2 ** 31, 2 ** 32 - 2])
def test_first_mb_out_of_range(first_mb):
    """first_mb_in_slice is ue(v) (up to 2^32 - 2): beyond the picture it must be refused, not used as an index."""
    data = sps() + pps() + idr_slice(first_mb, 1) + idr_slice(0, 12)
    frames, pics, bad = _decode("h264", data)
    assert bad == 0 and pics <= 1


@pytest.mark.parametrize("kw", [dict(log2_fn=13), dict(log2_fn=40), dict(log2_poc=13), dict(log2_poc=2 ** 20), dict(poc_type=3),
                                dict(w_mbs=2 ** 16), dict(h_mbs=2 ** 20), dict(w_mbs=1024, h_mbs=1024), dict(refs=17), dict(sps_id=32)])
def test_sps_out_of_range_is_ignored(kw):
    """An SPS with a field outside 7.4.2.1.1's range is dropped: the slices that follow find no parameter set."""
    data = sps(**kw) + pps(sps_id=kw.get("sps_id", 0) & 31) + idr_slice(0, 12)
    assert _decode("h264", data) == (0, 0, 0)


def test_slice_qp_out_of_range():
    for dq in (26, 30, -27, -100):
        assert _decode("h264", sps() + pps() + idr_slice(0, 12, qp_delta=dq))[1:] == (0, 0)
    assert _decode("h264", sps() + pps(init_qp=40) + idr_slice(0, 12))[1:] == (0, 0)


def test_size_change_without_new_frames_is_refused():
    """A second SPS of another size: the application's header callback re-sizes its buffers (set_frames), and
    pictures are decoded with the new geometry; a slice whose PPS points at an SPS of a size the frames were
    NOT set up for is refused."""
    a = sps(4, 3, sps_id=0) + pps(0, 0) + idr_slice(0, 12)
    b = sps(6, 5, sps_id=1) + pps(1, 1)
    # active SPS (last parsed, frames set for 6x5) is id 1; the slice uses PPS 0 -> SPS 0 (4x3)
    frames, pics, bad = _decode("h264", a + b + idr_slice(0, 12))
    assert bad == 0 and pics == 1


DAMAGED = sorted(f for f in os.listdir(os.path.join(util.GOLDEN, "damaged")) if f.endswith(".264"))


def test_damaged_mpeg2_without_sequence_header():
    """A sequence_extension with no sequence_header before it (width 8192, height 0): nothing may be decoded."""
    data = open(os.path.join(util.GOLDEN, "damaged", "mpeg2_22_3275.m2v"), "rb").read()
    assert _decode("mpeg2", data)[1:] == (0, 0)


@pytest.mark.parametrize("name", DAMAGED)
def test_damaged_fixtures(name):
    """Streams a fuzzing run of round 2 found: mb_qp_delta far out of range (a QP of 114 in the records), and
    slices running into macroblocks an earlier slice had delivered (stale work items)."""
    data = open(os.path.join(util.GOLDEN, "damaged", name), "rb").read()
    for chunk in (0, 13, 777):
        assert _decode("h264", data, chunk=chunk)[2] == 0


@pytest.mark.gpu
@pytest.mark.parametrize("name", DAMAGED)
def test_damaged_fixtures_on_the_device(name):
    """The same through the CUDA path: an error return or pictures, and the device stays usable afterwards."""
    data = open(os.path.join(util.GOLDEN, "damaged", name), "rb").read()
    try:
        m2dec_b200.decode_file("h264", data)
    except RuntimeError:
        pass
    good = open(os.path.join(util.GOLDEN, "h264_basic.264"), "rb").read()
    want = open(os.path.join(util.GOLDEN, "h264_basic.md5")).read().split()
    assert m2dec_b200.decode_file("h264", good) == want


# ---- the product path's records (strengths and, for Baseline streams, vectors left to the device) -------------------
def test_h264_damaged_streams_in_device_mode_records_are_sound():
    """With B2_PARSE_AS_DEVICE the front end emits what it sends to a device session when nobody looks on
    (bs_on_device, and mv_on_device with partitions instead of motion records for Baseline streams): those records
    must pass the structural check on damaged input as well.  The switch is read once per process: a child runs it."""
    import subprocess
    import sys
    code = r'''
import random, sys
sys.path.insert(0, %r); sys.path.insert(0, %r)
import test_robustness as T
for name in ("h264_basic", "h264_intra_ipcm", "h264_cabac_b_mixed"):
    data = T._golden(name, ".264")
    full = T._decode("h264", data)
    assert full[1] > 0 and full[2] == 0, (name, full)
    rng = random.Random(7)
    for cut in sorted(rng.sample(range(1, len(data)), 25)):
        assert T._decode("h264", data[:cut])[2] == 0, (name, cut)
    for trial in range(40):
        d = bytearray(data)
        for _ in range(rng.choice([1, 1, 2, 8, 64])):
            pos = rng.randrange(len(d))
            d[pos] = rng.randrange(256) if rng.random() < 0.5 else d[pos] ^ (1 << rng.randrange(8))
        assert T._decode("h264", bytes(d))[2] == 0, (name, trial)
print("ok")
''' % (os.path.dirname(os.path.dirname(os.path.abspath(__file__))), os.path.dirname(os.path.abspath(__file__)))
    env = dict(os.environ, B2_PARSE_AS_DEVICE="1")
    r = subprocess.run([sys.executable, "-c", code], env=env, stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, timeout=600)
    assert r.returncode == 0 and r.stdout.strip().endswith("ok"), r.stderr[-2000:]


@pytest.mark.gpu
@pytest.mark.parametrize("name", ["h264_basic", "h264_intra_ipcm", "h264_cabac_b_mixed"])
def test_h264_damaged_streams_through_the_device(name):
    """Damaged streams through the product path (device-derived strengths and vectors): every decode ends, and the
    device is still sound afterwards -- the clean stream decodes to its golden MD5s."""
    data = _golden(name, ".264")
    golden = open(os.path.join(util.GOLDEN, name + ".md5")).read().split()
    rng = random.Random(11)
    for trial in range(30):
        d = bytearray(data)
        if trial % 3 == 0:
            d = d[:rng.randrange(1, len(d))]
        else:
            for _ in range(rng.choice([1, 2, 8, 64])):
                pos = rng.randrange(len(d))
                d[pos] = rng.randrange(256) if rng.random() < 0.5 else d[pos] ^ (1 << rng.randrange(8))
        try:
            m2dec_b200.decode_file("h264", bytes(d))
        except RuntimeError:
            pass
    assert m2dec_b200.decode_file("h264", data) == golden
