#!/usr/bin/env python
"""bench.py — headline benchmark of m2dec-b200.

Metric (BASELINE.json): 1080p frames/sec/GPU, every frame bit-exact against the reference
decoder, on the H.264 Baseline CAVLC I/P 1920x1080 workload (BASELINE.json configs[1]):
`lanes` independent synthetic streams per GPU from the committed test-only writer.

  value  kernel replay: every picture's SoA resident in HBM, one batched launch per kernel and
         step (picture #t of all lanes); device-timed with CUDA events on the launching stream.
  e2e    the same metric through the reference-facing plugin table (m2d_func_table_t) with
         HOST buffers: elementary streams in host memory -> host parse -> H2D of the SoA ->
         kernels -> D2H of every output frame, timed end to end.
  roofline      dominant kernel: algorithmic bytes / its device time.
  cpu_baseline  the compiled reference CLI (oracle/_ref/h264dec, decode only) on the host cores.

`--impl reference` times the reference's own CPU decoder on the same streams, all host cores.
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import tempfile
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

STREAMGEN = os.path.join(ROOT, "tools", "_bin", "streamgen")
REF_BIN = os.path.join(ROOT, "oracle", "_ref", "h264dec")
E2EBENCH = os.path.join(ROOT, "tools", "_bin", "e2ebench")
# `value` and `e2e` are the same unit on the same streams, measured at two places (the docstring above): the name says so
METRIC = ("1080p frames/sec/GPU, H.264 Baseline CAVLC I/P, every frame bit-exact (MD5) vs the reference decoder -- "
          "value: reconstruction (pictures pre-parsed, records resident in HBM; parse, H2D and D2H excluded); "
          "e2e: elementary stream in host memory to frames in host memory through m2d_func_table_t")
UNIT = "frames/s"
MBW, MBH = 120, 68                      # 1920x1088 coded, 1920x1080 displayed
FRAME_BYTES = MBW * 16 * MBH * 16 * 3 // 2


def workload_args(seed, frames):
    # SURVEY.md 8d "C2": IDR + P, 1 reference, all P partition sizes, ~10 % intra, qp ~ 28,
    # deblocking on, alpha = beta = 0, bounded levels (hazards 1-2)
    return ("mbw=%d mbh=%d frames=%d seed=%d crop_bottom=8 gop=1000 refs=1 qp=28 intra_pct=10 skip_pct=15 "
            "coded_pct=25 deblock=1" % (MBW, MBH, frames, seed)).split()


def make_stream(seed, frames):
    return subprocess.run([STREAMGEN, "h264", *workload_args(seed, frames)], stdout=subprocess.PIPE, check=True).stdout


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


class ClockSampler:
    """SM clock and throttle reasons DURING the timed region (B200_PROFILING.md recipe).  NVML from a thread of this
    process (a sample every 20 ms; begin() / end() mark the timed region, samples outside it are dropped) -- an
    `nvidia-smi -lms` child needs the better part of a second to come up, more than a short timed region lasts;
    it remains the fallback where the NVML binding is not usable."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu):
        self.gpu, self.proc, self.path = gpu, None, None
        self.thread, self.stop_flag, self.samples, self.t0, self.t1 = None, False, [], None, None
        self.how = None

    def _nvml_loop(self, nv, h):
        reasons_fn = getattr(nv, "nvmlDeviceGetCurrentClocksEventReasons", None) or nv.nvmlDeviceGetCurrentClocksThrottleReasons
        while not self.stop_flag:
            try:
                self.samples.append((time.perf_counter(), nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM), int(reasons_fn(h))))
            except Exception:
                pass
            time.sleep(0.02)

    def start(self):
        try:
            import pynvml as nv
            nv.nvmlInit()
            # torch's device index follows CUDA_VISIBLE_DEVICES; NVML's does not
            vis = os.environ.get("CUDA_VISIBLE_DEVICES", "")
            idx = self.gpu
            if vis:
                ids = [v.strip() for v in vis.split(",") if v.strip()]
                if self.gpu < len(ids) and ids[self.gpu].isdigit():
                    idx = int(ids[self.gpu])
            h = nv.nvmlDeviceGetHandleByIndex(idx)
            self.nv, self.max_sm = nv, float(nv.nvmlDeviceGetMaxClockInfo(h, nv.NVML_CLOCK_SM))
            import threading
            self.thread = threading.Thread(target=self._nvml_loop, args=(nv, h), daemon=True)
            self.thread.start()
            self.how = "nvml"
            return
        except Exception:
            self.thread = None
        try:
            f = tempfile.NamedTemporaryFile(delete=False, suffix=".csv")
            self.path = f.name
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.gpu), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "100"], stdout=f, stderr=subprocess.DEVNULL)
            self.how = "nvidia-smi"
        except OSError:
            self.proc = None

    def begin(self):
        self.t0 = time.perf_counter()

    def end(self):
        self.t1 = time.perf_counter()

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": []}
        if self.thread:
            self.stop_flag = True
            self.thread.join(timeout=2)
            nv = self.nv
            inside = [x for x in self.samples if self.t0 is not None and self.t0 <= x[0] <= (self.t1 or 1e30)]
            names = [("hw_slowdown", "HwSlowdown"), ("hw_thermal_slowdown", "HwThermalSlowdown"),
                     ("sw_thermal_slowdown", "SwThermalSlowdown"), ("sw_power_cap", "SwPowerCap")]
            reasons = set()
            for _, _, r in inside:
                for n, suffix in names:
                    bit = getattr(nv, "nvmlClocksEventReason" + suffix, None) or getattr(nv, "nvmlClocksThrottleReason" + suffix, 0)
                    if r & bit:
                        reasons.add(n)
            if inside:
                out = {"sm_mhz": statistics.median([x[1] for x in inside]), "sm_max_mhz": self.max_sm, "reasons": sorted(reasons),
                       "samples": len(inside), "source": "nvml, every 20 ms inside the timed region"}
            return out
        if not self.proc:
            return out
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in open(self.path):
            f = [x.strip() for x in line.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1]))
                mx.append(float(f[2]))
            except ValueError:
                continue
            for n, v in zip(names, f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        os.unlink(self.path)
        if sm:
            out = {"sm_mhz": statistics.median(sm), "sm_max_mhz": max(mx), "reasons": sorted(reasons), "samples": len(sm),
                   "source": "nvidia-smi -lms 100 (started before the warm-up)"}
        return out


def run_reference_cli(paths, procs, md5=False):
    """Decode every file in `paths` with the compiled reference, `procs` processes at a time.
    Returns wall seconds."""
    t0 = time.perf_counter()
    running = []
    todo = list(paths)
    while todo or running:
        while todo and len(running) < procs:
            p = todo.pop()
            args = [REF_BIN] + (["-O"] if md5 else []) + [p]
            running.append(subprocess.Popen(args, cwd=os.path.dirname(p), stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL))
        running = [r for r in running if r.poll() is None]
        if running and (len(running) >= procs or not todo):
            time.sleep(0.002)
    return time.perf_counter() - t0


def run_oracle_port(streams, threads):
    """The CPU baseline when the compiled reference is not there: the product's host parse + the plain-C
    restatement of the reconstruction (oracle/, test infrastructure), `threads` streams at a time.
    Returns (wall seconds, MD5 lists)."""
    import util
    from concurrent.futures import ThreadPoolExecutor
    t0 = time.perf_counter()
    with ThreadPoolExecutor(max(1, threads)) as pool:
        md5s = list(pool.map(lambda s: util.oracle_md5("h264", s), streams))
    return time.perf_counter() - t0, md5s


def write_streams(streams, d):
    paths = []
    for i, s in enumerate(streams):
        p = os.path.join(d, "s%03d.264" % i)
        with open(p, "wb") as f:
            f.write(s)
        paths.append(p)
    return paths


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--lanes", type=int, default=512, help="independent 1080p streams per GPU")
    ap.add_argument("--frames", type=int, default=16, help="pictures per stream (IDR + P)")
    ap.add_argument("--prep-threads", type=int, default=16, help="host threads that write and parse the lanes' streams (untimed)")
    ap.add_argument("--e2e-streams", type=int, default=32, help="streams per end-to-end repetition (at least; sixteen per host thread are used when that is more)")
    ap.add_argument("--e2e-reps", type=int, default=6, help="timed repetitions of the end-to-end pool run")
    ap.add_argument("--verify-lanes", type=int, default=32, help="lanes whose every frame is MD5-checked against the reference")
    a = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    ncores = os.cpu_count() or 1
    workload = "h264_baseline_cavlc_1080p_ip"

    if a.impl == "reference":
        if rank != 0:
            return 0
        procs = max(1, min(ncores, a.lanes))
        streams = [make_stream(100 + i, a.frames) for i in range(procs)]
        kind = "reference" if os.path.exists(REF_BIN) else "port"
        if kind == "reference":
            with tempfile.TemporaryDirectory() as d:
                paths = write_streams(streams, d)
                for _ in range(a.warmup):
                    run_reference_cli(paths, procs)
                t = 0.0
                for _ in range(a.steps):
                    t += run_reference_cli(paths, procs)
            sample = "%d streams x %d frames per step, %d processes of the compiled reference CLI (decode only)" % (procs, a.frames, procs)
        else:
            # oracle/_ref was not built (no /root/reference at build time): the CPU restatement stands in
            for _ in range(min(a.warmup, 1)):
                run_oracle_port(streams, procs)
            t = 0.0
            for _ in range(a.steps):
                t += run_oracle_port(streams, procs)[0]
            sample = "%d streams x %d frames per step, %d threads of the oracle port (host parse + plain-C reconstruction)" % (procs, a.frames, procs)
        frames = procs * a.frames * a.steps
        v = frames / t
        print(json.dumps({
            "impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": a.gpus, "steps": a.steps,
            "warmup": a.warmup, "ms_per_step": 1000 * t / a.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "u8", "data": "synthetic",
            "config": {"workload": workload, "lanes_per_gpu": a.lanes, "frames_per_lane": a.frames, "coded_size": "1920x1088",
                       "display_size": "1920x1080", "sample": sample},
            "cpu_baseline": {"value": v, "unit": UNIT, "cores": procs, "kind": kind, "sample": sample},
            "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        }))
        return 0

    # stdout carries exactly one JSON line: anything a library prints there meanwhile (NCCL's version
    # banner, for one) is sent to stderr
    json_fd = os.dup(1)
    os.dup2(2, 1)

    import numpy as np
    import torch
    import m2dec_b200
    from m2dec_b200 import _native as L
    from m2dec_b200 import shard

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device - the reconstruction path has no CPU fallback")
    torch.cuda.set_device(local)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group(backend="nccl", device_id=torch.device("cuda", local))

    # ---- build the workload: parse on the host once, make every picture resident in HBM ------
    t_prep = time.perf_counter()
    seeds = shard.rank_seeds(rank, a.lanes)
    rp = m2dec_b200.Replay("h264", MBW * 16, MBH * 16, a.lanes, pool_frames=18, cuda_device=local)
    streams = [None] * a.lanes

    def prepare(lane):
        streams[lane] = make_stream(seeds[lane], a.frames)
        n = rp.load_stream(lane, streams[lane])
        assert n == a.frames, (n, a.frames)

    if a.prep_threads > 1:
        from concurrent.futures import ThreadPoolExecutor
        with ThreadPoolExecutor(max(1, min(ncores // world, a.prep_threads))) as pool:
            list(pool.map(prepare, range(a.lanes)))
    else:
        for lane in range(a.lanes):
            prepare(lane)
    t_prep = time.perf_counter() - t_prep

    # ---- parity (untimed): every frame of the first lanes against the compiled reference -----
    rp.run()
    rp.sync()
    verified, verify_kind = 0, "none"
    goldens = {}
    nver = min(a.verify_lanes, a.lanes)
    if not os.path.exists(REF_BIN) and nver:
        # no compiled reference on this machine: the oracle port (pinned to the reference by tests/) is the checker
        import hashlib
        nver = min(nver, 2)
        verify_kind = "oracle port (oracle/recon_h264.c)"
        for lane, md5s in enumerate(run_oracle_port(streams[:nver], nver)[1]):
            goldens[lane] = md5s
            for t, idx in enumerate(rp.outputs[lane]):
                y, c = rp.download(lane, idx)
                h = hashlib.md5()
                h.update(np.ascontiguousarray(y[:1080]).tobytes())
                h.update(np.ascontiguousarray(c[:540]).tobytes())
                if h.hexdigest() != md5s[t]:
                    raise SystemExit("bench.py: MD5 mismatch lane %d picture %d" % (lane, t))
                verified += 1
    with tempfile.TemporaryDirectory() as d:
        if os.path.exists(REF_BIN) and nver:
            paths = write_streams(streams[:nver], d)
            run_reference_cli(paths, min(ncores, nver), md5=True)
            verify_kind = "compiled reference (oracle/_ref/h264dec -O), hashed on the device (b2_replay_md5)"
            import hashlib
            # every picture of the verified lanes hashed on the device in one launch (b2_replay_md5: 16 bytes per picture
            # come back); the first two lanes are also downloaded and hashed on the host, which checks the device MD5 itself
            all_got = rp.md5([(lane, idx) for lane in range(nver) for idx in rp.outputs[lane]], width=1920, height=1080)
            pos = 0
            for lane in range(nver):
                golden = open(os.path.join(d, "s%03d.out" % lane), "rb").read().split()
                if len(golden) != a.frames:
                    raise SystemExit("bench.py: the reference decoded %d of %d frames of lane %d" % (len(golden), a.frames, lane))
                goldens[lane] = [g.decode() for g in golden]
                got = all_got[pos:pos + len(rp.outputs[lane])]
                pos += len(rp.outputs[lane])
                if [g.encode() for g in got] != golden:
                    raise SystemExit("bench.py: MD5 mismatch in lane %d" % lane)
                for t, idx in enumerate(rp.outputs[lane] if lane < 2 else []):
                    y, c = rp.download(lane, idx)
                    h = hashlib.md5()
                    h.update(np.ascontiguousarray(y[:1080]).tobytes())
                    h.update(np.ascontiguousarray(c[:540]).tobytes())
                    if h.hexdigest().encode() != golden[t]:
                        raise SystemExit("bench.py: MD5 mismatch lane %d picture %d" % (lane, t))
                verified += len(got)

    # ---- timed region: K steps, one step = all lanes x frames ---------------------------------
    def barrier():
        if dist is not None:
            dist.barrier()
        rp.sync()
        torch.cuda.synchronize()

    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    for _ in range(a.warmup):
        rp.run()
    barrier()
    kernel_ms = [0.0, 0.0, 0.0]
    sampler.begin()
    rp.mark(0)
    for _ in range(a.steps):
        rp.run(timed=True)
    rp.mark(1)
    rp.sync()
    barrier()
    sampler.end()
    total_ms = rp.elapsed_ms()
    st = rp.stats()
    clocks = sampler.stop() if rank == 0 else None
    # per-kernel device time of the last timed step (events between the phases of every sub-step)
    kernel_ms = [st.kernel_ms[i] for i in range(3)]
    frames_per_step = a.lanes * a.frames
    value = shard.whole_job_rate(frames_per_step * a.steps, total_ms / 1000.0, dist, "cuda")
    total_ms = shard.max_over_ranks(total_ms, dist, "cuda")

    # ---- end to end through the plugin table with host buffers --------------------------------
    threads_n = max(1, min(a.lanes, ncores // world))     # the ranks of a node share its host cores
    # streams per repetition: at least sixteen per host thread, so that a repetition measures the pool's rate and not the
    # start-up and drain of one short stream per thread (a 32-core box ran the former 32 streams one per thread)
    ne = min(a.lanes, max(a.e2e_streams, 16 * threads_n))
    threads_n = min(threads_n, ne)
    counts = [None] * ne
    frames_out = [0] * ne

    e2e_checked = [0]

    def decode_one(i, check=False):
        n = [0]
        md5s = []

        def post(frm):
            n[0] += 1
            if check:
                md5s.append(m2dec_b200.frame_md5(frm))

        dec = m2dec_b200.M2Decoder("h264", streams[i], cuda_device=local)
        try:
            dec.run(post)
            c = (L.C.c_uint64 * 3)()
            L.lib.b2_context_stats(dec.ctx, L.C.byref(c))
            counts[i] = [int(c[0]), int(c[1]), int(c[2])]
        finally:
            dec.close()
        frames_out[i] = n[0]
        if check:
            if md5s != goldens[i]:
                raise SystemExit("bench.py: e2e MD5 mismatch on stream %d" % i)
            e2e_checked[0] += len(md5s)

    def run_pool(check=False):
        work = list(range(ne))
        lock = threading.Lock()

        def worker():
            while True:
                with lock:
                    if not work:
                        return
                    i = work.pop()
                decode_one(i, check and i in goldens)

        ths = [threading.Thread(target=worker) for _ in range(threads_n)]
        for t in ths:
            t.start()
        for t in ths:
            t.join()

    e2e_digest = None
    run_pool(check=True)   # the Python host (m2dec_b200.M2Decoder): untimed, every frame MD5-checked against the reference
    e2e_host = "m2dec_b200.M2Decoder (Python threads)"
    native = os.path.exists(E2EBENCH) and not os.environ.get("B2_BENCH_PYTHON_E2E")
    if native:
        # The timed host is native (tools/e2ebench: C++ threads around the same plugin table, as the reference
        # application is C++): the Python threads cost 20 % of the end-to-end rate at 16 threads.
        e2e_host = "tools/e2ebench (C++ threads)"
        with tempfile.TemporaryDirectory() as d:
            paths = write_streams(streams[:ne], d)
            # untimed: every frame's MD5 (computed by e2ebench on the frames in host memory) against the reference's lines
            r = subprocess.run([E2EBENCH, "h264", str(threads_n), "1", str(local), "1"] + paths, stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True)
            if r.returncode != 0:
                raise SystemExit("bench.py: e2ebench failed: " + r.stderr[-500:])
            for line in r.stdout.splitlines():
                f = line.split()
                if f and f[0] == "stream":
                    i = int(f[1])
                    if int(f[2]) != a.frames:
                        raise SystemExit("bench.py: e2ebench stream %d gave %s frames" % (i, f[2]))
                    if i in goldens:
                        if f[3:] != goldens[i]:
                            raise SystemExit("bench.py: e2ebench MD5 mismatch on stream %d" % i)
                        e2e_checked[0] += len(goldens[i])
            barrier()
            # the first three repetitions are the warm-up (CUDA context, page-locked frame buffers, the library's buffer cache: the
            # second pass still allocates where its threads meet the streams in another order)
            WARM = 3
            r = subprocess.run([E2EBENCH, "h264", str(threads_n), str(a.e2e_reps + WARM), str(local), "0"] + paths, stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True)
            if r.returncode != 0:
                raise SystemExit("bench.py: e2ebench failed: " + r.stderr[-500:])
            reps = [float(l.split()[-2]) for l in r.stderr.splitlines() if l.startswith("e2ebench repetition")]
            assert len(reps) == a.e2e_reps + WARM, r.stderr
            for t_rep in reps[WARM:]:
                sys.stderr.write("e2e repetition: %d frames in %.3f s\n" % (ne * a.frames, t_rep))
            tot = r.stdout.split()
            e2e_frames = ne * a.frames * a.e2e_reps
            e2e_s = sum(reps[WARM:])
            h2d, d2h = int(tot[tot.index("h2d") + 1]), int(tot[tot.index("d2h") + 1])
            # Digest mode (SURVEY 8f rank 1 behind the plugin table, what `h264dec -O` produces): frames hashed on the
            # device, 16 bytes per frame come back.  Same streams and threads; every digest of the verified streams is
            # compared with the reference's line.  Reported next to e2e, not instead of it.
            if world == 1 and not os.environ.get("B2_BENCH_NO_DIGEST"):
                r = subprocess.run([E2EBENCH, "h264", str(threads_n), str(a.e2e_reps + WARM), str(local), "2"] + paths, stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True)
                if r.returncode != 0:
                    raise SystemExit("bench.py: e2ebench (digest mode) failed: " + r.stderr[-500:])
                dchecked = 0
                for line in r.stdout.splitlines():
                    f = line.split()
                    if f and f[0] == "stream" and int(f[1]) in goldens:
                        if f[3:] != goldens[int(f[1])]:
                            raise SystemExit("bench.py: device MD5 mismatch on stream %s (digest mode)" % f[1])
                        dchecked += len(f) - 3
                dreps = [float(l.split()[-2]) for l in r.stderr.splitlines() if l.startswith("e2ebench repetition")]
                dtot = r.stdout.split()
                e2e_digest = {"value": ne * a.frames * a.e2e_reps / sum(dreps[WARM:]), "unit": UNIT,
                              "h2d_bytes_per_step": int(dtot[dtot.index("h2d") + 1]), "d2h_bytes_per_step": int(dtot[dtot.index("d2h") + 1]),
                              "md5_checked_frames": dchecked,
                              "path": "m2d_func_table_t + b2_context_set_md5 / b2_context_frame_md5: host ES -> parse -> H2D SoA -> kernels -> MD5 of the cropped frame on the device -> 16 bytes D2H per frame"}
    else:
        run_pool()             # second warm-up pass in the timed configuration (no MD5 on the worker threads)
        barrier()
        e2e_frames = 0
        t0 = time.perf_counter()
        for _ in range(a.e2e_reps):
            t_rep = time.perf_counter()
            run_pool()
            e2e_frames += sum(frames_out)
            sys.stderr.write("e2e repetition: %d frames in %.3f s\n" % (sum(frames_out), time.perf_counter() - t_rep))
        torch.cuda.synchronize()
        e2e_s = time.perf_counter() - t0
        h2d = sum(c[0] for c in counts)      # bytes of one repetition (= one pass over the e2e streams)
        d2h = sum(c[1] for c in counts)
    e2e_value = shard.whole_job_rate(e2e_frames, e2e_s, dist, "cuda")
    e2e_s = shard.max_over_ranks(e2e_s, dist, "cuda")

    if rank != 0:
        if dist is not None:
            dist.destroy_process_group()
        return 0

    # ---- roofline of the dominant kernel ------------------------------------------------------
    peak, peak_src = peaks()
    # the kernels behind the three phases of a replay step (m2dec_b200/csrc/h264_recon.cu)
    names = ["h264_inter5_kernel", "h264_intra_sparse_kernel (P) / h264_wave_kernel<1> (I)", "h264_deblock3_kernel"]
    phase = ["inter", "intra", "deblock"]
    dom = max(range(3), key=lambda i: kernel_ms[i])
    alg = [int(st.alg_bytes[i]) for i in range(3)]
    achieved = alg[dom] / (kernel_ms[dom] / 1000.0) / 1e9 if kernel_ms[dom] > 0 else 0.0
    whole = sum(alg) / (st.total_ms / 1000.0) / 1e9 if st.total_ms > 0 else 0.0
    # DRAM bytes of one launch of the dominant kernel from the committed ncu --set full capture,
    # scaled to this run's pictures per launch
    traffic = None
    tpath = os.path.join(ROOT, "profiles", "r02_traffic.json")       # written by tools/summarize_ncu.py from the round's ncu --set full capture
    if os.path.exists(tpath):
        tj = json.load(open(tpath))
        per_launch = tj["dram_bytes_per_launch"].get(phase[dom])
        if per_launch:
            traffic = per_launch / tj["pictures_per_launch"] * a.lanes
    roofline = {"bound": "hbm", "kernel": names[dom], "achieved": achieved, "peak": peak, "unit": "GB/s",
                "frac": achieved / peak, "traffic": traffic, "traffic_unit": "bytes per launch (ncu dram__bytes_read+write, P-picture step)",
                "alg_bytes_per_launch": alg[dom] / max(1, a.frames - (1 if dom == 0 else 0)),
                "peak_source": peak_src,
                "per_kernel": {names[i]: {"ms_per_step": kernel_ms[i], "alg_bytes_per_step": alg[i],
                                          "GBps": (alg[i] / (kernel_ms[i] / 1000.0) / 1e9 if kernel_ms[i] > 0 else None)}
                               for i in range(3)},
                "whole_step": {"alg_bytes": sum(alg), "ms": st.total_ms, "GBps": whole, "frac": whole / peak}}

    # ---- CPU baseline: compiled reference on a bounded sample (about 10-30 s of CPU work) -----
    cpu = None
    if not os.path.exists(REF_BIN):
        procs = max(1, min(ncores, a.lanes, 32))
        tcpu, _ = run_oracle_port(streams[:procs], procs)
        cpu = {"value": procs * a.frames / tcpu, "unit": UNIT, "cores": procs, "kind": "port",
               "sample": "%d of the workload's streams x %d frames, oracle port (host parse + plain-C reconstruction), %d threads, %d host cores present"
                         % (procs, a.frames, procs, ncores)}
    if os.path.exists(REF_BIN):
        procs = max(1, min(ncores, a.lanes, 32))
        nsample = min(a.lanes, 3 * procs)            # about 10-20 s of CPU work at 16 frames per stream
        with tempfile.TemporaryDirectory() as d:
            paths = write_streams(streams[:nsample], d)
            run_reference_cli(paths[:1], 1)
            tcpu = run_reference_cli(paths, procs)
        cpu = {"value": nsample * a.frames / tcpu, "unit": UNIT, "cores": procs, "kind": "reference",
               "sample": "%d of the workload's streams x %d frames, one reference CLI process per stream (decode only), %d at a time, %d host cores present"
                         % (nsample, a.frames, procs, ncores)}

    out = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": a.steps, "warmup": a.warmup,
        "ms_per_step": total_ms / a.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "u8", "data": "synthetic",
        "config": {"workload": workload, "lanes_per_gpu": a.lanes, "frames_per_lane": a.frames,
                   "frames_per_step": frames_per_step, "coded_size": "1920x1088", "display_size": "1920x1080",
                   "l2_policy": "inputs larger than L2 (%.0f MB of pictures written per step)" % (frames_per_step * FRAME_BYTES / 1e6),
                   "parity": "%d frames MD5-checked against %s before timing" % (verified, verify_kind),
                   "partitioning": "independent streams per GPU, no collective", "host_prep_s": t_prep},
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                "streams": ne, "repetitions": a.e2e_reps, "frames": e2e_frames, "md5_checked_frames": e2e_checked[0], "host_threads": threads_n, "seconds": e2e_s,
                "host": e2e_host,
                "path": "m2d_func_table_t (h264d_gpu_func): host ES -> entropy decode -> H2D SoA (partitions with vector differences, coefficient levels) -> kernels (motion vector prediction, work list, strengths, reconstruction) -> D2H every output frame"},
        "gpu_launches": int(st.launches) * a.steps,
        "roofline": roofline,
        "cpu_baseline": cpu,
        "clocks": clocks,
    }
    if e2e_digest is not None:
        out["e2e_device_md5"] = e2e_digest
    sys.stdout.flush()
    os.write(json_fd, (json.dumps(out) + "\n").encode())
    if dist is not None:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
