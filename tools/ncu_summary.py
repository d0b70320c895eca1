#!/usr/bin/env python
"""Condense ncu output (run here, no GPU needed) into small tracked files under profiles/.
  python tools/ncu_summary.py launches gpurun_out/launches.csv profiles/r01_launches.md
  python tools/ncu_summary.py full gpurun_out/prof_x.ncu-rep profiles/r01_x.md
"""
import collections, csv, io, re, subprocess, sys

def launches(src, dst):
    rows = list(csv.reader(open(src, errors="ignore")))
    # find header
    hi = next(i for i, r in enumerate(rows) if "Kernel Name" in r)
    hdr = rows[hi]; ci = {h: i for i, h in enumerate(hdr)}
    agg = collections.OrderedDict()
    for r in rows[hi + 1:]:
        if len(r) < len(hdr) or r[ci["Metric Name"]] != "gpu__time_duration.sum":
            continue
        name = re.sub(r"\(.*", "", r[ci["Kernel Name"]]).replace("void ", "").replace("saunet::", "")
        v = float(r[ci["Metric Value"]].replace(",", "")); unit = r[ci["Metric Unit"]]
        us = v / 1e3 if unit in ("ns", "nsecond") else (v if unit in ("us", "usecond") else v * 1e3)
        e = agg.setdefault(name, [0, 0.0]); e[0] += 1; e[1] += us
    tot = sum(e[1] for e in agg.values())
    with open(dst, "w") as f:
        f.write("# ncu launch list (gpu__time_duration.sum, --clock-control none; cold-cache, serialised: compare SHARES)\n\n")
        f.write("source: `%s`; %d launches, %.2f ms total\n\n| kernel | launches | total ms | share |\n|---|---:|---:|---:|\n" % (src, sum(e[0] for e in agg.values()), tot / 1e3))
        for k, e in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            f.write("| `%s` | %d | %.3f | %.1f%% |\n" % (k, e[0], e[1] / 1e3, 100 * e[1] / tot))
    print("wrote", dst)

KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct",
        "smsp__inst_executed.sum", "sm__cycles_elapsed.avg", "launch__grid_size", "launch__block_size",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio", "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio"]

def full(src, dst):
    out = subprocess.run(["ncu", "-i", src, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units = rows[0], rows[1]
    with open(dst, "w") as f:
        f.write("# ncu --set full summary of `%s`\n\n" % src)
        for r in rows[2:]:
            d = dict(zip(hdr, r))
            f.write("## %s\n\n| metric | value | unit |\n|---|---:|---|\n" % d.get("Kernel Name", "?"))
            for k in KEYS:
                if k in d:
                    f.write("| %s | %s | %s |\n" % (k, d[k], units[hdr.index(k)]))
            f.write("\n")
        sass = subprocess.run(["ncu", "-i", src, "--page", "source", "--csv"], capture_output=True, text=True).stdout
        ops = collections.Counter()
        for r in csv.reader(io.StringIO(sass)):
            if len(r) > 5 and r[0].startswith("0x"):
                t = r[1].split(); op = (t[1] if t and t[0].startswith("@") and len(t) > 1 else (t[0] if t else "?")).split(".")[0]
                try: ops[op] += int(r[5])
                except ValueError: pass
        tot = sum(ops.values()) or 1
        f.write("Executed warp instructions by opcode (top 12): " + ", ".join("%s %.1f%%" % (k, 100 * v / tot) for k, v in ops.most_common(12)) + "\n\n")
        f.write("Blackwell evidence in SASS: " + ", ".join("%s x%d" % (k, v) for k, v in ops.items() if k.startswith(("UTC", "LDTM", "UBLKCP", "UTMA"))) + "\n")
    print("wrote", dst)

if __name__ == "__main__":
    {"launches": launches, "full": full}[sys.argv[1]](sys.argv[2], sys.argv[3])
