"""How much of an end-to-end repetition tools/e2ebench's threads had work for (the rest is the tail of the pool run);
run on the GPU box: python tools/e2e_util.py"""
import os, subprocess, sys, tempfile
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); import bench
from concurrent.futures import ThreadPoolExecutor
with ThreadPoolExecutor(16) as ex: data = list(ex.map(lambda i: bench.make_stream(100 + i, 16), range(128)))
d = tempfile.mkdtemp(); paths = bench.write_streams(data, d)
r = subprocess.run([bench.E2EBENCH, 'h264', '16', '4', '0', '0'] + paths, stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True)
print(r.stderr[-600:])
