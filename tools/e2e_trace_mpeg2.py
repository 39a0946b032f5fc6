"""MPEG-2 (BASELINE config 4) end to end: rate at 16 and 1 host threads, the parse alone, the reference CLI on one stream;
run on the GPU box: python tools/e2e_trace_mpeg2.py"""
import os, subprocess, sys, tempfile, re, collections
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); import bench
from concurrent.futures import ThreadPoolExecutor
args = "mbw=120 mbh=68 frames=12 seed=%d bframes=2 gop=12 mv_range=64"
with ThreadPoolExecutor(16) as ex: data = list(ex.map(lambda i: subprocess.run([bench.STREAMGEN, 'mpeg2', *(args % (100+i)).split()], stdout=subprocess.PIPE, check=True).stdout, range(64)))
print('stream bytes', len(data[0]), 'per frame', len(data[0])//12)
d = tempfile.mkdtemp(); paths = []
for i, s in enumerate(data):
    p = os.path.join(d, 's%03d.m2v' % i); open(p, 'wb').write(s + b'\xff'*32); paths.append(p)
for th in (16, 1):
    n = 64 if th > 1 else 8
    r = subprocess.run([bench.E2EBENCH, 'mpeg2', str(th), '3', '0', '0'] + paths[:n], stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, env=dict(os.environ, B2_TRACE='1'))
    miss = collections.Counter(l for l in r.stderr.splitlines() if 'buffer cache miss' in l)
    tl = [l for l in r.stderr.splitlines() if l.startswith('m2dec_b200 trace') and 'per picture:' in l][-n:]
    acc = collections.defaultdict(float)
    for l in tl:
        for name, v in re.findall(r' ([a-zA-Z0-9 ()]+?) ([0-9.]+),', l.split('per picture:')[1]): acc[name.strip()] += float(v)
    reps = [float(l.split()[-2]) for l in r.stderr.splitlines() if l.startswith('e2ebench repetition')]
    print('threads', th, 'fps', n * 12 / reps[-1], 'ms/frame/thread', 1000 * reps[-1] * th / (n * 12), {k: round(v / max(1,len(tl)), 1) for k, v in acc.items()}, 'buffer cache misses', dict(miss), 'repetitions', reps)
# parse alone
r = subprocess.run([os.path.join(bench.ROOT, 'tools/_bin/parsebench'), 'mpeg2', paths[0], '20'], stdout=subprocess.PIPE, text=True); print(r.stdout.strip())
import time
t0=time.perf_counter(); subprocess.run([bench.REF_BIN, paths[0]], cwd=d, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL); print('reference CLI one stream', (time.perf_counter()-t0)*1000/12, 'ms per frame incl. process start')
