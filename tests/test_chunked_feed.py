"""The application may hand the stream over in pieces through one reused buffer (reference
src/app: dec_bits_set_callback + dec_bits_set_data from the refill callback; bytes of an earlier
refill are gone).  Units and start codes then straddle buffers at every possible offset: the
decoded pictures must not depend on the chunking.  Product parse + CPU oracle, no GPU."""
import os

import pytest

import streams
import util

CHUNKS = [1, 2, 3, 5, 64, 1000, 4099]


def _golden(name, ext):
    data = open(os.path.join(util.GOLDEN, name + ext), "rb").read()
    md5 = open(os.path.join(util.GOLDEN, name + ".md5")).read().split()
    return data, md5


@pytest.mark.parametrize("name", streams.H264_GOLDEN)
def test_h264_chunked_feed(name):
    data, golden = _golden(name, ".264")
    for chunk in CHUNKS:
        assert util.oracle_md5("h264", data, chunk=chunk) == golden, chunk


@pytest.mark.parametrize("name", streams.MPEG2_GOLDEN + ["cv2_mpeg2_352x288"])
def test_mpeg2_chunked_feed(name):
    data, golden = _golden(name, ".m2v")
    for chunk in CHUNKS:
        assert util.oracle_md5("mpeg2", data, chunk=chunk) == golden, chunk


def test_four_byte_start_codes_and_padding():
    """Zero bytes before start codes (00 00 00 01, padding) in every chunking."""
    data, golden = _golden("h264_basic", ".264")
    padded = data.replace(b"\x00\x00\x01", b"\x00\x00\x00\x00\x01")
    assert padded != data
    for chunk in [0] + CHUNKS:
        assert util.oracle_md5("h264", padded, chunk=chunk) == golden, chunk
