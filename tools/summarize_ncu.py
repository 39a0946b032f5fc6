"""Turns the ncu captures of a round into the tracked summaries under profiles/.

  python tools/summarize_ncu.py full   <report.ncu-rep> <pictures per launch> <out.txt> <traffic.json>
  python tools/summarize_ncu.py launch <launches.csv> <out.txt>

`full`: one block per captured kernel launch of an `ncu --set full` report (duration, registers,
occupancy, issue rate, instruction counts, DRAM bytes, cache hit rates, the stall reasons above 0.3
warps per issue) and the DRAM bytes per launch of each reconstruction phase (bench.py scales them
into roofline.traffic).  `launch`: the `--metrics gpu__time_duration.sum` launch list of the bench
command, aggregated per kernel and grid."""
import collections
import csv
import json
import re
import subprocess
import sys

PHASE = [("h264_inter5_kernel", "inter"), ("h264_intra_sparse_kernel", "intra"), ("h264_deblock3_kernel", "deblock")]
KEEP = ["launch__grid_size", "launch__block_size", "launch__registers_per_thread", "gpu__time_duration.sum",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "smsp__inst_executed.sum", "smsp__thread_inst_executed_per_inst_executed.ratio", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active"]


def short(name):
    m = re.search(r"(h264_\w+|mpeg2_\w+|md5_\w+)(<[^>]*>)?", name)
    return (m.group(1) + (m.group(2) or "")) if m else name


def full(rep, pics, out_txt, out_json):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], stdout=subprocess.PIPE, text=True, check=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units = rows[0], rows[1]
    col = {h: i for i, h in enumerate(hdr)}
    lines = ["ncu --set full --clock-control none --import-source on, B200; %d pictures of 1920x1088 per launch; report %s (scratch, not tracked)" % (pics, rep), ""]
    traffic = {}
    for r in rows[2:]:
        name = short(r[col["Kernel Name"]])
        lines.append("== %s   grid %s block %s" % (name, r[col["Grid Size"]] if "Grid Size" in col else "", r[col["Block Size"]] if "Block Size" in col else ""))
        for k in KEEP:
            if k in col:
                lines.append("  %-80s %s %s" % (k, r[col[k]], units[col[k]]))
        for h, i in col.items():
            if "issue_stalled" in h and h.endswith("per_issue_active.ratio"):
                try:
                    v = float(r[i])
                except ValueError:
                    continue
                if v >= 0.3:
                    lines.append("  %-80s %.2f" % (h.replace("smsp__average_warps_issue_stalled_", "stall ").replace("_per_issue_active.ratio", ""), v))
        try:
            rd, wr = float(r[col["dram__bytes_read.sum"]]), float(r[col["dram__bytes_write.sum"]])
            scale = {"Mbyte": 1e6, "Gbyte": 1e9, "Kbyte": 1e3, "byte": 1.0}[units[col["dram__bytes_read.sum"]]]
            scale_w = {"Mbyte": 1e6, "Gbyte": 1e9, "Kbyte": 1e3, "byte": 1.0}[units[col["dram__bytes_write.sum"]]]
            total = rd * scale + wr * scale_w
            inst = float(r[col["smsp__inst_executed.sum"]])
            lines.append("  %-80s %.1f MB  (%.2f MB per picture)" % ("dram traffic per launch", total / 1e6, total / 1e6 / pics))
            lines.append("  %-80s %.0f" % ("warp instructions per macroblock (8160 MB x %d pictures)" % pics, inst / (8160.0 * pics)))
            for pat, label in PHASE:
                if pat in r[col["Kernel Name"]]:
                    traffic[label] = total          # the last launch of a phase wins (a P-picture step)
        except (ValueError, KeyError):
            pass
        lines.append("")
    open(out_txt, "w").write("\n".join(lines))
    json.dump({"pictures_per_launch": pics, "dram_bytes_per_launch": traffic, "source": rep}, open(out_json, "w"), indent=1)


def launch(path, out_txt):
    rows = [r for r in csv.reader(open(path)) if len(r) > 14 and r[0].isdigit()]
    agg = collections.OrderedDict()
    for r in rows:
        key = (short(r[4]), r[8])
        a = agg.setdefault(key, [0, 0.0])
        a[0] += 1
        a[1] += float(r[14]) / 1e6
    def is_replay(k, g):       # batched launches: the strip kernel has one grid row per picture, the persistent kernels many CTAs
        x, y, _ = [int(t) for t in g.strip("()").split(",")]
        if k.startswith(("h264_mv", "h264_items", "h264_bs")):       # derivation kernels: the plugin path's (one picture per launch)
            return False
        return y > 1 or (not k.startswith("h264_inter5") and not k.startswith("mpeg2") and x > 64)
    big = sum(v[1] for (k, g), v in agg.items() if is_replay(k, g))
    lines = ["kernel                                 grid             launches     total ms     avg us    share of the replay launches"]
    for (k, g), v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        replay = is_replay(k, g)
        label = ("%.1f%%" % (100 * v[1] / big)) if replay else ("verification (untimed)" if "md5" in k else "e2e")
        lines.append("%-38s %-16s %8d %12.3f %10.1f %8s" % (k[:38], g, v[0], v[1], 1000 * v[1] / v[0], label))
    open(out_txt, "w").write("\n".join(lines) + "\n")


if __name__ == "__main__":
    if sys.argv[1] == "full":
        full(sys.argv[2], int(sys.argv[3]), sys.argv[4], sys.argv[5])
    else:
        launch(sys.argv[2], sys.argv[3])
