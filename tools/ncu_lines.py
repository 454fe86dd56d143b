"""Summarise an `ncu --page source --csv --print-source cuda,sass` dump per source line:
python tools/ncu_lines.py dump.csv [top_n] -> share of executed instructions, thread efficiency, no-instruction stall samples."""
import csv
import sys
from collections import defaultdict

path = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
hdr = None
cur = None
rows = []
for r in csv.reader(open(path)):
    if len(r) == 2 and r[0] == "File Path":
        cur = r[1].split("/")[-1]
        continue
    if r and r[0] == "Line No":
        hdr = r
        continue
    if hdr is None or len(r) != len(hdr) or r[0] == "":
        continue
    try:
        ln = int(r[0])
        inst = int(r[hdr.index("Instructions Executed")])
        thr = int(r[hdr.index("Thread Instructions Executed")])
        samples = int(r[hdr.index("# Samples")])
        noinst = int(r[hdr.index("stall_no_inst")])
        longsb = int(r[hdr.index("stall_long_sb")])
    except ValueError:
        continue
    rows.append((cur, ln, r[1].strip()[:80], inst, thr, samples, noinst, longsb))
ti = sum(x[3] for x in rows)
ts = sum(x[5] for x in rows)
tn = sum(x[6] for x in rows)
print(f"instructions {ti:.3e}  avg threads {sum(x[4] for x in rows) / ti:.1f}  samples {ts}  no_inst share of samples {tn / ts:.2%}")
byfile = defaultdict(lambda: [0, 0, 0, 0])
for x in rows:
    b = byfile[x[0]]
    b[0] += x[3]; b[1] += x[4]; b[2] += x[5]; b[3] += x[6]
for k, v in byfile.items():
    print(f"  {k:20s} inst {v[0] / ti:6.1%}  thr/inst {v[1] / max(v[0], 1):5.1f}  samples {v[2] / ts:6.1%}  no_inst {v[3] / max(tn, 1):6.1%}")
rows.sort(key=lambda x: -x[5])
print("top lines by stall samples:")
for x in rows[:top]:
    print(f"  {x[0]:18s}{x[1]:5d} samp {x[5] / ts:5.1%} inst {x[3] / ti:5.1%} thr {x[4] / max(x[3], 1):4.1f} noinst {x[6] / max(x[5], 1):4.0%} longsb {x[7] / max(x[5], 1):4.0%}  {x[2]}")
