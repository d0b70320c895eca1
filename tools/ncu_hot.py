#!/usr/bin/env python
"""Hottest SASS instructions (stall samples + executed counts) of an .ncu-rep captured with --import-source on.
   python tools/ncu_hot.py file.ncu-rep [top]"""
import csv, io, subprocess, sys
out = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "source", "--csv"], capture_output=True, text=True).stdout
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
rows = list(csv.reader(io.StringIO(out)))
hdr = rows[1]
ia, isrc, isamp, iexec = hdr.index("Address"), hdr.index("Source"), hdr.index("# Samples"), hdr.index("Instructions Executed")
data = []
for n, r in enumerate(rows[2:]):
    try: data.append((n, r[isrc].strip(), int(r[isamp]), int(r[iexec])))
    except (ValueError, IndexError): pass
ts, te = sum(d[2] for d in data), sum(d[3] for d in data)
print("total samples %d, executed warp-instr %d, SASS lines %d" % (ts, te, len(data)))
print("--- by stall samples")
for n, s, sm, ex in sorted(data, key=lambda d: -d[2])[:top]:
    print("%5d %6.2f%% samp %6.2f%% exec  %s" % (n, 100.0 * sm / ts, 100.0 * ex / te, s[:110]))
print("--- by executed")
for n, s, sm, ex in sorted(data, key=lambda d: -d[3])[:top // 2]:
    print("%5d %6.2f%% samp %6.2f%% exec  %s" % (n, 100.0 * sm / ts, 100.0 * ex / te, s[:110]))
