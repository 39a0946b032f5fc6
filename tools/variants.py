"""Kernel experiments: builds variants of the library with other -D flags into m2dec_b200/variants/ (git-ignored,
they travel to the GPU box) and, with --run, times each with tools/profile_replay.py (B2_LIB selects the build).
    python tools/variants.py build name1:-DX=1 name2:-DX=2,-DY=3
    python tools/variants.py run [--lanes 256 --frames 4] name1 name2      (on the GPU box)
"""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "m2dec_b200"))
VDIR = os.path.join(ROOT, "m2dec_b200", "variants")


def lib(name):
    return os.path.join(VDIR, "lib_%s.so" % name)


if sys.argv[1] == "build":
    import build as B
    os.makedirs(VDIR, exist_ok=True)
    for spec in sys.argv[2:]:
        name, _, flags = spec.partition(":")
        extra = [f for f in flags.split(",") if f]
        print(name, extra, B.build(force=True, out=lib(name), extra=extra, objdir=os.path.join("/tmp/b2_variants", name)), flush=True)
elif sys.argv[1] == "run":
    args = sys.argv[2:]
    opts = []
    while args and args[0].startswith("--"):
        opts += args[:2]
        args = args[2:]
    for name in args:
        env = dict(os.environ, B2_LIB=lib(name))
        r = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "profile_replay.py")] + opts, env=env, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
        print("==", name, "rc", r.returncode)
        print(r.stdout.strip()[-600:], flush=True)
