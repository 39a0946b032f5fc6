"""Breaks the end-to-end path down: parse only (device off) vs full decode, N host threads."""
import os, sys, threading, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench, m2dec_b200

N = int(sys.argv[1]) if len(sys.argv) > 1 else 16
F = int(sys.argv[2]) if len(sys.argv) > 2 else 16
streams = [bench.make_stream(100 + i, F) for i in range(N)]

def run(device_off, threads):
    work = list(range(N)); lock = threading.Lock(); frames = [0]
    def one(i):
        dec = m2dec_b200.M2Decoder("h264", streams[i], device_off=device_off)
        n = [0]
        try:
            dec.run(lambda f: n.__setitem__(0, n[0] + 1))
        finally:
            dec.close()
        with lock: frames[0] += n[0]
    def worker():
        while True:
            with lock:
                if not work: return
                i = work.pop()
            one(i)
    ths = [threading.Thread(target=worker) for _ in range(threads)]
    t0 = time.perf_counter()
    for t in ths: t.start()
    for t in ths: t.join()
    dt = time.perf_counter() - t0
    return frames[0] / dt, dt

TH = [int(x) for x in sys.argv[3].split(",")] if len(sys.argv) > 3 else [1, 4, 8, 16]
OFF = (False,) if os.environ.get("PROBE_DEVICE_ONLY") else (True, False)
for off in OFF:
    for th in TH:
        if th > N: continue
        run(off, th)
        fps, dt = run(off, th)
        print("device_off=%s threads=%2d  %.1f fps  (%.1f ms/frame/thread)" % (off, th, fps, 1000 * th / fps))
