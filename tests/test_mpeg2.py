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
