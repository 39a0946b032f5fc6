# Where a host thread's time goes in the end-to-end leg (B2_TRACE=1 lines of the library, last repetition), at 16, 8 and 1
# threads; run on the GPU box: bash tools/e2e_trace.sh
cd "$(dirname "$0")/.."
python - <<'PY'
import sys, os, subprocess, tempfile, re, collections
sys.path.insert(0, os.getcwd()); import bench
from concurrent.futures import ThreadPoolExecutor
with ThreadPoolExecutor(16) as ex: data = list(ex.map(lambda i: bench.make_stream(100 + i, 16), range(64)))
d = tempfile.mkdtemp(); paths = bench.write_streams(data, d)
for th in (16, 8, 1):
    env = dict(os.environ, B2_TRACE='1')
    n = 64 if th > 1 else 8
    r = subprocess.run([bench.E2EBENCH, 'h264', str(th), '3', '0', '0'] + paths[:n], stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, env=env)
    acc = collections.defaultdict(float); cnt = 0
    tl = [l for l in r.stderr.splitlines() if l.startswith('m2dec_b200 trace')][-n:]      # the last repetition's sessions
    for l in tl:
        if l.startswith('m2dec_b200 trace'):
            cnt += 1
            for name, v in re.findall(r' ([a-zA-Z0-9 ()]+?) ([0-9.]+),', l.split('per picture:')[1]): acc[name.strip()] += float(v)
    reps = [float(l.split()[-2]) for l in r.stderr.splitlines() if l.startswith('e2ebench repetition')]
    print('threads', th, 'fps', n * 16 / reps[-1], 'ms/frame/thread', 1000 * reps[-1] * th / (n * 16), {k: round(v / cnt, 1) for k, v in acc.items()}, 'sum', round(sum(acc.values()) / cnt, 1))
PY
