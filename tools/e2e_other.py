"""End-to-end rate of the configurations that are parity cases, not bench lines (BASELINE configs 3 and 4), next to the
reference CLI on the same host cores:  python tools/e2e_other.py   (on the GPU box)"""
import os, subprocess, sys, tempfile
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
from concurrent.futures import ThreadPoolExecutor

threads = os.cpu_count() or 1
REPS, WARM = 4, 2        # (the first repetition allocates the sessions' buffers, the second still finds some caches cold)
CASES = {
    "c3 h264 High CABAC I/P/B 1080p": ("h264", "264", "h264", "mbw=120 mbh=68 frames=16 seed=%d crop_bottom=8 refs=2 bframes=2 cabac=1 profile=100 transform8x8=1 weighted_pred=1 weighted_bipred=1 gop=1000 qp=28 intra_pct=10 skip_pct=15 coded_pct=25", b"", 8),
    "c4 mpeg2 MP@HL I/P/B 1080p": ("mpeg2", "m2v", "mpeg2", "mbw=120 mbh=68 frames=12 seed=%d bframes=2 gop=12 mv_range=64", b"\xff" * 32, 16),
}
for name, (codec, ext, gen, args, tail, per_thread) in CASES.items():
    N = per_thread * threads           # streams per repetition (a repetition must be long next to its tail: >= 0.3 s)
    with tempfile.TemporaryDirectory() as d:
        def make(i):
            return subprocess.run([bench.STREAMGEN, gen, *(args % (100 + i)).split()], stdout=subprocess.PIPE, check=True).stdout
        with ThreadPoolExecutor(threads) as ex:
            data = list(ex.map(make, range(N)))
        paths = []
        for i, s in enumerate(data):
            p = os.path.join(d, "s%03d.%s" % (i, ext))
            open(p, "wb").write(s + tail)          # (MPEG-2: the reference needs the padding to reach its end of stream)
            paths.append(p)
        frames = int(args.split("frames=")[1].split()[0])
        r = subprocess.run([bench.E2EBENCH, codec, str(threads), str(REPS + WARM), "0", "0"] + paths, stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True)
        t = [float(l.split()[-2]) for l in r.stderr.splitlines() if l.startswith("e2ebench repetition")]
        ours = N * frames * REPS / sum(t[WARM:]) if r.returncode == 0 and len(t) == REPS + WARM else float("nan")
        ref_s = bench.run_reference_cli(paths, threads) if os.path.exists(bench.REF_BIN) else float("nan")
        print("%-34s ours %8.1f frames/s   reference CLI %7.1f frames/s   (%d host threads, %d streams x %d frames)" % (name, ours, N * frames / ref_s, threads, N, frames), flush=True)
        if r.returncode != 0:
            print(r.stderr[-300:])
