"""Where the CPU time of an e2ebench_sample run went: by mapping (library), and by function inside libm2dec_b200.so."""
import collections
import subprocess
import sys

pcs = [int(x, 16) for x in open(sys.argv[1]).read().split()]
maps = []
for line in open(sys.argv[2]):
    f = line.split()
    if len(f) >= 6 and "x" in f[1]:
        a, b = [int(x, 16) for x in f[0].split("-")]
        maps.append((a, b, int(f[2], 16), f[5]))
by_map = collections.Counter()
ours = collections.Counter()
lib = None
for pc in pcs:
    for a, b, off, name in maps:
        if a <= pc < b:
            by_map[name.split("/")[-1]] += 1
            if "libm2dec_b200" in name:
                lib = name
                ours[pc - a + off] += 1
            break
    else:
        by_map["?"] += 1
total = len(pcs)
print("samples", total)
for k, v in by_map.most_common(12):
    print("%5.1f%%  %s" % (100.0 * v / total, k))
if lib and len(sys.argv) > 3:
    lib = sys.argv[3]
if lib:
    addrs = ["%x" % a for a in ours]
    out = subprocess.run(["addr2line", "-e", lib, "-f", "-C"] + addrs, stdout=subprocess.PIPE, text=True).stdout.split("\n")
    fn = collections.Counter()
    for i, a in enumerate(ours):
        fn[out[2 * i][:90]] += ours[a]
    print("\n== inside libm2dec_b200.so")
    for k, v in fn.most_common(25):
        print("%5.1f%%  %s" % (100.0 * v / total, k))
