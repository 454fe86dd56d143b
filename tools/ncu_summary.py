"""Per-kernel summary of an `ncu --set full ... ; ncu -i rep --page raw --csv` dump:
  python tools/ncu_summary.py raw.csv [--traffic-json profiles/ncu_traffic.json]
Prints one markdown row per kernel name (averaged over its captured launches): duration, registers, achieved occupancy, issue
utilisation, threads per instruction, executed instructions, DRAM read/write bytes, L1 hit rate for local loads, top stall
reasons (per issued instruction). With --traffic-json, writes {kernel short name: dram read+write bytes per launch} for bench.py."""
import argparse
import csv
import json
import re
from collections import defaultdict


def f(x):
    try:
        return float(x.replace(",", ""))
    except Exception:
        return None


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("csv")
    ap.add_argument("--traffic-json")
    args = ap.parse_args()
    rows = list(csv.reader(open(args.csv)))
    hdr, units = rows[0], rows[1]
    col = {h: i for i, h in enumerate(hdr)}

    def find(*subs):
        for h in hdr:
            if all(s in h for s in subs):
                return col[h]
        return None
    c_name = col["Kernel Name"]
    want = {
        "ms": find("gpu__time_duration.sum"), "regs": find("launch__registers_per_thread"), "occ": find("sm__warps_active.avg.pct_of_peak_sustained_active"),
        "issue": find("sm__inst_issued.avg.pct_of_peak_sustained_active") or find("smsp__issue_active.avg.pct"), "thr": find("smsp__thread_inst_executed_per_inst_executed.ratio"),
        "inst": find("smsp__inst_executed.sum"), "dram_r": find("dram__bytes_read.sum"), "dram_w": find("dram__bytes_write.sum"),
        "l1_local": find("l1tex__t_sector_pipe_lsu_mem_local_op_ld_hit_rate") or find("mem_local_op_ld", "hit_rate"),
    }
    stall_cols = {h.split("smsp__average_warps_issue_stalled_")[1].split("_per_issue_active")[0]: i for h, i in col.items() if "smsp__average_warps_issue_stalled_" in h and "per_issue_active" in h and "not_issued" not in h}
    unit_scale = {"Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12, "byte": 1.0, "us": 1e-3, "ms": 1.0, "s": 1e3, "ns": 1e-6, "usecond": 1e-3, "msecond": 1.0, "second": 1e3, "nsecond": 1e-6}
    agg = defaultdict(lambda: defaultdict(list))
    for r in rows[2:]:
        if len(r) != len(hdr):
            continue
        name = re.sub(r"\(.*", "", r[c_name]).replace("void ", "").strip()
        name = re.sub(r"^.*::", "", name) if "k_tsvq" in name or "<unnamed>" in name else name
        if name in ("k_candidates",):
            name += " grid " + r[col["Grid Size"]].replace(" ", "")   # one row per work-list launch (LA / RGB / alpha classes)
        for k, i in want.items():
            if i is None:
                continue
            v = f(r[i])
            if v is None:
                continue
            agg[name][k].append(v * unit_scale.get(units[i], 1.0))
        for k, i in stall_cols.items():
            v = f(r[i])
            if v is not None:
                agg[name]["stall_" + k].append(v)
    print("| kernel | launches | ms | regs | warps active % | issue % | threads/inst | instructions | DRAM rd + wr | L1 hit local ld % | top stalls per issue |")
    print("|---|---|---|---|---|---|---|---|---|---|---|")
    traffic = {}
    for name, d in agg.items():
        m = {k: sum(v) / len(v) for k, v in d.items()}
        n = len(d.get("ms", [0]))
        stalls = sorted(((k[6:], v) for k, v in m.items() if k.startswith("stall_")), key=lambda kv: -kv[1])[:4]
        tr = (m.get("dram_r", 0) + m.get("dram_w", 0))
        traffic[name] = tr
        print(f"| {name} | {n} | {m.get('ms', 0):.3f} | {m.get('regs', 0):.0f} | {m.get('occ', 0):.1f} | {m.get('issue', 0):.1f} | {m.get('thr', 0):.1f} | {m.get('inst', 0):.3g} | "
              f"{m.get('dram_r', 0) / 1e9:.3f} + {m.get('dram_w', 0) / 1e9:.3f} GB | {m.get('l1_local', float('nan')):.1f} | " + ", ".join(f"{k} {v:.1f}" for k, v in stalls) + " |")
    if args.traffic_json:
        old = {}
        try:
            old = json.load(open(args.traffic_json))
        except Exception:
            pass
        old.update({k: v for k, v in traffic.items()})
        cand = [v for k, v in traffic.items() if k.startswith("k_candidates")]
        if cand:
            old["k_candidates"] = sum(cand)   # bench.py reports the three work-list launches of a step together
        json.dump(old, open(args.traffic_json, "w"), indent=1, sort_keys=True)


if __name__ == "__main__":
    main()
