"""Generates tests/golden/cv2_mpeg2_*.m2v with OpenCV's bundled ffmpeg MPEG-2 encoder (an
encoder independent of both the reference and this repo) and the matching *.md5 files by
running the compiled reference decoder (oracle/_ref/h264dec -O).  Run in the build
container, where /root/reference exists:   python tests/golden/make_cv2_mpeg2.py
"""
import os
import sys

import cv2
import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
import util  # noqa: E402


def make(name, w, h, n, seed):
    path = os.path.join(HERE, name + ".m2v")
    wr = cv2.VideoWriter(path, cv2.VideoWriter_fourcc(*"mpg2"), 25, (w, h))
    rng = np.random.default_rng(seed)
    base = cv2.GaussianBlur(rng.integers(0, 255, (h + 4 * n, w + 6 * n, 3), dtype=np.uint8), (9, 9), 3)
    for i in range(n):
        f = base[i * 2:i * 2 + h, i * 3:i * 3 + w].copy()
        f[h // 4:h // 2, (7 * i) % (w // 2):(7 * i) % (w // 2) + w // 4] //= 2   # a moving dark box
        wr.write(np.ascontiguousarray(f))
    wr.release()
    data = open(path, "rb").read()
    md5 = util.ref_md5(data, "m2v")
    assert len(md5) == n, (len(md5), n)
    with open(os.path.join(HERE, name + ".md5"), "w") as f:
        f.write("\n".join(md5) + "\n")
    print(name, len(data), "bytes", n, "frames")


if __name__ == "__main__":
    make("cv2_mpeg2_352x288", 352, 288, 15, 0)
