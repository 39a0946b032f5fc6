"""Buffer-cache misses (cudaMalloc / cudaMallocHost inside a session) per repetition of tools/e2ebench, from the B2_TRACE
lines of the library; run on the GPU box: python tools/e2e_cache_misses.py"""
import os, subprocess, sys, tempfile
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); import bench
from concurrent.futures import ThreadPoolExecutor
with ThreadPoolExecutor(16) as ex: data = list(ex.map(lambda i: bench.make_stream(100 + i, 16), range(128)))
d = tempfile.mkdtemp(); paths = bench.write_streams(data, d)
r = subprocess.run([bench.E2EBENCH, 'h264', '16', '6', '0', '0'] + paths, stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, env=dict(os.environ, B2_TRACE='1'))
rep = 0; miss = {}
for l in r.stderr.splitlines():
    if 'cache miss' in l: miss.setdefault(rep, []).append(l.split('miss: ')[1])
    if l.startswith('e2ebench repetition'): print(l, '| misses in it:', len(miss.get(rep, [])), sorted(set(miss.get(rep, [])))[:6]); rep += 1
