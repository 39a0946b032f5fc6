"""Builds libm2dec_b200.so in-tree with nvcc for sm_100a (no torch involved: the library is a
plain C-ABI shared object, see include/*.h)."""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
CSRC = os.path.join(HERE, "csrc")
OUT = os.path.join(HERE, "libm2dec_b200.so")

CPP = ["bits.cpp", "backend.cpp", "plugin.cpp", "validate.cpp", "mpeg2_front.cpp", "h264_front.cpp"]
CU = ["session.cu", "mpeg2_recon.cu", "h264_recon.cu", "replay.cu", "md5.cu"]

NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]


def _newer(src_list, out):
    if not os.path.exists(out):
        return True
    t = os.path.getmtime(out)
    deps = list(src_list)
    for d in (CSRC, os.path.join(ROOT, "include")):
        deps += [os.path.join(d, f) for f in os.listdir(d)]
    return any(os.path.getmtime(p) > t for p in deps if os.path.isfile(p))


def _host_has_v3():
    try:
        flags = next(l for l in open("/proc/cpuinfo") if l.startswith("flags")).split()
    except (OSError, StopIteration):
        return False
    return all(f in flags for f in ("avx2", "bmi1", "bmi2", "abm", "fma", "movbe"))


def build(force=False, verbose=False, out=None, extra=(), objdir=None):
    """out / extra / objdir: a variant build (other -D flags) next to the product library, for kernel experiments."""
    OUT = out or globals()["OUT"]
    srcs = [os.path.join(CSRC, f) for f in CPP + CU if os.path.exists(os.path.join(CSRC, f))]
    if not force and not _newer(srcs, OUT):
        return OUT
    objdir = objdir or os.path.join(HERE, "build")
    os.makedirs(objdir, exist_ok=True)
    common = ["-O3", "-std=c++17", "-lineinfo", "-I" + os.path.join(ROOT, "include"), "-I" + CSRC,
              "-Xcompiler", "-fPIC,-Wall,-Wno-unused-function"]
    # host parse (the end-to-end bottleneck): x86-64-v3 (AVX2, BMI2, LZCNT -- every x86 host a B200 board sits in) gives
    # the entropy decoders three-operand shifts and lzcnt: 2.5 % (CAVLC) to 4 % (MPEG-2) of the parse
    host_arch = ["-Xcompiler", "-march=x86-64-v3"] if _host_has_v3() else []
    procs = []
    objs = []
    for s in srcs:
        o = os.path.join(objdir, os.path.basename(s) + ".o")
        objs.append(o)
        cmd = [NVCC] + ARCH + common + os.environ.get("B2_NVCC_EXTRA", "").split() + list(extra) + (host_arch if s.endswith(".cpp") else []) + (["-Xptxas", "-v"] if verbose else []) + ["-c", s, "-o", o]
        procs.append((cmd, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
    failed = False
    for cmd, p in procs:
        out, _ = p.communicate()
        if p.returncode != 0 or verbose:
            sys.stderr.write(" ".join(cmd) + "\n" + out + "\n")
        failed |= p.returncode != 0
    if failed:
        raise RuntimeError("nvcc failed")
    cmd = [NVCC] + ARCH + ["-shared", "-o", OUT] + objs + ["-Xcompiler", "-fPIC", "-lpthread"]
    r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if r.returncode != 0:
        sys.stderr.write(r.stdout)
        raise RuntimeError("link failed")
    return OUT


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
