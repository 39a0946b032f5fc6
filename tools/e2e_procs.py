"""End-to-end decode with P processes instead of P threads (is the Python host the bottleneck?)."""
import os, sys, time, subprocess
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
if len(sys.argv) > 1 and sys.argv[1] == "worker":
    import bench, m2dec_b200
    ids = [int(x) for x in sys.argv[2].split(",")]
    streams = [bench.make_stream(100 + i, 16) for i in ids]
    def one(s):
        dec = m2dec_b200.M2Decoder("h264", s)
        n = [0]
        try:
            dec.run(lambda f: n.__setitem__(0, n[0] + 1))
        finally:
            dec.close()
        return n[0]
    one(streams[0])                      # warm-up
    print("ready", flush=True)
    sys.stdin.readline()
    t0 = time.perf_counter()
    frames = sum(one(s) for s in streams for _ in range(3))
    print(frames, time.perf_counter() - t0, flush=True)
else:
    P = int(sys.argv[1]) if len(sys.argv) > 1 else 16
    procs = [subprocess.Popen([sys.executable, __file__, "worker", "%d,%d" % (2 * i, 2 * i + 1)], stdin=subprocess.PIPE, stdout=subprocess.PIPE, text=True) for i in range(P)]
    for p in procs:
        assert p.stdout.readline().strip() == "ready"
    t0 = time.perf_counter()
    for p in procs:
        p.stdin.write("go\n"); p.stdin.flush()
    frames = 0
    for p in procs:
        f, dt = p.stdout.readline().split()
        frames += int(f)
    dt = time.perf_counter() - t0
    print("processes=%d  %d frames in %.3f s = %.1f fps" % (P, frames, dt, frames / dt))
