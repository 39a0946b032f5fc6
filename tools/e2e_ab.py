"""A/B of the end-to-end leg (tools/e2ebench) under different environments, same streams, on the GPU box:
    python tools/e2e_ab.py [streams] [reps] -- NAME=ENV1=1,ENV2=x NAME2= ...
prints frames/s of the timed repetitions (the first two are the warm-up) per configuration, two rounds."""
import os, subprocess, sys, tempfile
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench

args = sys.argv[1:]
cut = args.index("--") if "--" in args else len(args)
nums = [int(x) for x in args[:cut]]
n_streams = nums[0] if nums else 64
reps = nums[1] if len(nums) > 1 else 5
configs = args[cut + 1:] or ["default="]
threads = int(os.environ.get("E2E_THREADS", os.cpu_count() or 1))
with tempfile.TemporaryDirectory() as d:
    from concurrent.futures import ThreadPoolExecutor
    with ThreadPoolExecutor(threads) as ex:
        data = list(ex.map(lambda i: bench.make_stream(100 + i, 16), range(n_streams)))
    paths = bench.write_streams(data, d)
    for rnd in range(2):
        for c in configs:
            name, _, envs = c.partition("=")
            env = dict(os.environ)
            for kv in envs.split(","):
                if kv:
                    k, _, v = kv.partition("=")
                    env[k] = v or "1"
            r = subprocess.run([bench.E2EBENCH, "h264", str(threads), str(reps + 2), "0", "0"] + paths, stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, env=env)
            t = [float(l.split()[-2]) for l in r.stderr.splitlines() if l.startswith("e2ebench repetition")]
            if r.returncode != 0 or len(t) != reps + 2:
                print(name, "FAILED", r.stderr[-300:])
                continue
            print("%-12s %8.1f frames/s  (threads %d, %d streams, repetitions %s)" % (name, n_streams * 16 * reps / sum(t[2:]), threads, n_streams, " ".join("%.3f" % x for x in t[2:])), flush=True)
