cd $GRAFT_REPO_ROOT
python - <<'PY'
import os, sys, threading, time
sys.path.insert(0, '.')
sys.argv=['x','16','16']
if os.environ.get('WITH_TORCH'):
    import torch; torch.cuda.set_device(0); torch.cuda.synchronize()
import bench, m2dec_b200
N=16; F=16
streams = [bench.make_stream(100 + i, F) for i in range(N)]
def run(threads):
    work = list(range(N)); lock = threading.Lock(); frames = [0]
    def one(i):
        dec = m2dec_b200.M2Decoder("h264", streams[i])
        n = [0]
        try: dec.run(lambda f: n.__setitem__(0, n[0] + 1))
        finally: dec.close()
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
    return frames[0] / dt
run(16)
print('torch' if os.environ.get('WITH_TORCH') else 'plain', [round(run(16)) for _ in range(4)])
PY
