"""Short kernel-replay run for profiling (ncu): parse a few synthetic 1080p streams, make the SoA
resident, run the replay a few times.  No e2e, no verification (tests/ and bench.py do that).
--codec mpeg2 runs BASELINE config 4 (MPEG-2 MP@HL 1920x1088 I/P/B) through mpeg2_recon_kernel."""
import argparse
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
import m2dec_b200  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--codec", default="h264", choices=["h264", "mpeg2"])
ap.add_argument("--workload", default="c2", choices=["c2", "c3"], help="H.264: c2 = Baseline CAVLC I/P (the bench workload), c3 = High CABAC I/P/B with weighted prediction and the 8x8 transform (BASELINE config 3)")
ap.add_argument("--lanes", type=int, default=16)
ap.add_argument("--frames", type=int, default=4)
ap.add_argument("--runs", type=int, default=2)
ap.add_argument("--distinct", type=int, default=0, help="number of distinct streams (0 = one per lane); the others are repeats: shorter preparation for kernel A/B runs")
ap.add_argument("--threads", type=int, default=os.cpu_count() or 1, help="host threads that write and load the streams (1 under ncu)")
a = ap.parse_args()


def stream(seed):
    if a.codec == "h264" and a.workload == "c3":
        args = ("mbw=%d mbh=%d frames=%d seed=%d crop_bottom=8 refs=2 bframes=2 cabac=1 profile=100 transform8x8=1 "
                "weighted_pred=1 weighted_bipred=1 gop=1000 qp=28 intra_pct=10 skip_pct=15 coded_pct=25" % (bench.MBW, bench.MBH, a.frames, seed))
        return subprocess.run([bench.STREAMGEN, "h264", *args.split()], stdout=subprocess.PIPE, check=True).stdout
    if a.codec == "h264":
        return bench.make_stream(seed, a.frames)
    args = "mbw=%d mbh=%d frames=%d seed=%d bframes=2 gop=12 mv_range=64" % (bench.MBW, bench.MBH, a.frames, seed)
    return subprocess.run([bench.STREAMGEN, "mpeg2", *args.split()], stdout=subprocess.PIPE, check=True).stdout


rp = m2dec_b200.Replay(a.codec, bench.MBW * 16, bench.MBH * 16, a.lanes, pool_frames=18 if a.codec == "h264" else 4)
from concurrent.futures import ThreadPoolExecutor
seeds = sorted({100 + (lane % a.distinct if a.distinct else lane) for lane in range(a.lanes)})
with ThreadPoolExecutor(a.threads) as ex:
    _cache = dict(zip(seeds, ex.map(stream, seeds)))


def _load(lane):
    n = rp.load_stream(lane, _cache[100 + (lane % a.distinct if a.distinct else lane)])
    assert n == a.frames, (n, a.frames)


with ThreadPoolExecutor(a.threads) as ex:
    list(ex.map(_load, range(a.lanes)))
for _ in range(a.runs):
    rp.run(timed=True)
    rp.sync()
st = rp.stats()
alg = [int(st.alg_bytes[i]) for i in range(3)]
ms = [st.kernel_ms[i] for i in range(3)]
print("pictures", st.pictures, "launches", st.launches, "ms", ms, st.total_ms)
print("alg GB/s per kernel", [round(alg[i] / ms[i] / 1e6, 1) if ms[i] > 0 else None for i in range(3)],
      "fps", round(st.pictures / st.total_ms * 1000.0, 1))
rp.close()
