"""Per-function and per-line shares of the program counters parsebench_sample wrote (pcs.txt)."""
import collections
import subprocess
import sys

exe, pcs = sys.argv[1], sys.argv[2]
addrs = open(pcs).read().split()
cnt = collections.Counter(addrs)
uniq = list(cnt)
out = subprocess.run(["addr2line", "-e", exe, "-f", "-C", "-i"] + uniq, stdout=subprocess.PIPE, text=True).stdout.split("\n")
# addr2line -i prints (function, file:line) pairs, innermost first, for every address; re-run one by one is slow,
# so use -a to separate
out = subprocess.run(["addr2line", "-e", exe, "-f", "-C", "-i", "-a"] + uniq, stdout=subprocess.PIPE, text=True).stdout.split("\n")
frames = {}
cur = None
i = 0
while i < len(out):
    l = out[i]
    if l.startswith("0x"):
        cur = l[2:].lstrip("0")
        frames[cur] = []
        i += 1
        continue
    if cur is not None and i + 1 < len(out) and l:
        frames[cur].append((l, out[i + 1].split("/")[-1].split(" ")[0]))
        i += 2
        continue
    i += 1
total = sum(cnt.values())
by_line, by_outer, by_inner = collections.Counter(), collections.Counter(), collections.Counter()
for a, n in cnt.items():
    fr = frames.get(a.lstrip("0"), [])
    if not fr:
        by_outer["?"] += n
        continue
    by_inner[fr[0][0][:70]] += n
    by_outer[fr[-1][0][:70]] += n
    by_line[fr[0][1] + "  <" + fr[-1][1] + ">"] += n
print("samples", total)
print("\n== by outermost (non-inlined) function")
for k, v in by_outer.most_common(25):
    print("%5.1f%%  %s" % (100.0 * v / total, k))
print("\n== by innermost (inlined) function")
for k, v in by_inner.most_common(40):
    print("%5.1f%%  %s" % (100.0 * v / total, k))
print("\n== by line (innermost <outermost>)")
for k, v in by_line.most_common(70):
    print("%5.1f%%  %s" % (100.0 * v / total, k))
