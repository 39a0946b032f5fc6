"""Short kernel-replay run for profiling (ncu): parse a few synthetic 1080p streams, make the SoA
resident, run the replay a few times.  No e2e, no verification (tests/ and bench.py do that)."""
import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
import m2dec_b200  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--lanes", type=int, default=16)
ap.add_argument("--frames", type=int, default=4)
ap.add_argument("--runs", type=int, default=2)
a = ap.parse_args()
rp = m2dec_b200.Replay("h264", bench.MBW * 16, bench.MBH * 16, a.lanes, pool_frames=18)
for lane in range(a.lanes):
    rp.load_stream(lane, bench.make_stream(100 + lane, a.frames))
for _ in range(a.runs):
    rp.run(timed=True)
    rp.sync()
st = rp.stats()
print("pictures", st.pictures, "launches", st.launches, "ms", [st.kernel_ms[i] for i in range(3)], st.total_ms)
rp.close()
