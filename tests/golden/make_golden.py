"""Writes tests/golden/<name>.264|.m2v (test-only writer, fixed seeds) and <name>.md5 (MD5 lines
of the compiled reference decoder, oracle/_ref/h264dec -O).  Run in the build container:
    python tests/golden/make_golden.py
"""
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
import streams  # noqa: E402
import util  # noqa: E402

for name in streams.H264_GOLDEN:
    data = streams.generate("h264", streams.H264_CASES[name])
    md5 = util.ref_md5(data, "264")
    assert md5, name
    open(os.path.join(HERE, name + ".264"), "wb").write(data)
    open(os.path.join(HERE, name + ".md5"), "w").write("\n".join(md5) + "\n")
    print(name, len(data), "bytes", len(md5), "frames")
for name in getattr(streams, "MPEG2_GOLDEN", []):
    data = streams.generate("mpeg2", streams.MPEG2_CASES[name])
    md5 = util.ref_md5(data, "m2v")
    assert md5, name
    open(os.path.join(HERE, name + ".m2v"), "wb").write(data)
    open(os.path.join(HERE, name + ".md5"), "w").write("\n".join(md5) + "\n")
    print(name, len(data), "bytes", len(md5), "frames")
