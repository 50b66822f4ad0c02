"""Turns the ncu metric list of one 64-spp pass's traversal launches (tools/gpu_r02h.sh: every trace_closest / trace_shadow
launch, `--metrics dram bytes, lanes per instruction, issue utilisation ...`) into profiles/traversal_profile.json, the file
bench.py reads for roofline.traffic / dram_frac / issue.   python tools/summarise_traversal_profile.py <csv> <tag>"""
import collections
import csv
import json
import sys
from pathlib import Path

REPO = Path(__file__).resolve().parents[1]


def main():
    src, tag = Path(sys.argv[1]), sys.argv[2]
    global LAST
    LAST = int(sys.argv[3]) if len(sys.argv) > 3 else 10
    rows = [r for r in csv.reader(src.open()) if len(r) > 6]
    hdr = rows[0]
    col = {n: hdr.index(n) for n in ("ID", "Kernel Name", "Metric Name", "Metric Unit", "Metric Value")}
    scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}
    launches = collections.OrderedDict()
    for r in rows[1:]:
        kid = r[col["ID"]]
        kind = "closest" if "trace_closest" in r[col["Kernel Name"]] else "shadow"
        d = launches.setdefault(kid, {"kind": kind})
        try:
            v = float(r[col["Metric Value"]].replace(",", ""))
        except ValueError:
            continue
        d[r[col["Metric Name"]]] = v * scale.get(r[col["Metric Unit"]], 1.0)
    out = {"source": f"profiles/{tag}_traversal_metrics.csv (ncu --clock-control none, every traversal launch of one 64-spp pass of C3)"}
    for kind in ("closest", "shadow"):
        ls = [d for d in launches.values() if d["kind"] == kind and "gpu__time_duration.sum" in d]
        ls = ls[-LAST:]  # the capture starts inside the warm-up pass: keep the launches of the timed pass (one per bounce)
        n = len(ls)
        t = sum(d["gpu__time_duration.sum"] for d in ls)
        inst = sum(d["smsp__inst_executed.sum"] for d in ls)
        wavg = lambda key: sum(d[key] * d["gpu__time_duration.sum"] for d in ls) / t  # noqa: E731
        lanes = sum(d["smsp__thread_inst_executed_per_inst_executed.ratio"] * d["smsp__inst_executed.sum"] for d in ls) / inst
        summary = {
            "launches": n, "ms_total_under_ncu": round(t, 3),
            "dram_bytes_per_launch": int(sum(d["dram__bytes_read.sum"] + d["dram__bytes_write.sum"] for d in ls) / n),
            "dram_read_bytes_total": int(sum(d["dram__bytes_read.sum"] for d in ls)),
            "dram_write_bytes_total": int(sum(d["dram__bytes_write.sum"] for d in ls)),
            "lanes_per_instruction": round(lanes, 2),
            "issue_active_pct": round(wavg("smsp__issue_active.avg.pct_of_peak_sustained_active"), 2),
            "warps_active_pct": round(wavg("sm__warps_active.avg.pct_of_peak_sustained_active"), 2),
            "l1_hit_pct": round(wavg("l1tex__t_sector_hit_rate.pct"), 2), "l2_hit_pct": round(wavg("lts__t_sector_hit_rate.pct"), 2),
            "warp_instructions": int(inst),
            "per_launch": [{"ms": round(d["gpu__time_duration.sum"], 3), "lanes": d["smsp__thread_inst_executed_per_inst_executed.ratio"],
                            "issue_pct": round(d["smsp__issue_active.avg.pct_of_peak_sustained_active"], 1),
                            "dram_gb": round((d["dram__bytes_read.sum"] + d["dram__bytes_write.sum"]) * 1e-9, 3)} for d in ls],
        }
        out[kind] = summary
    c = out["closest"]
    out["dram_bytes_per_launch"] = c["dram_bytes_per_launch"]
    out["issue"] = {"issue_active_pct": c["issue_active_pct"], "lanes_per_instruction": c["lanes_per_instruction"],
                    "simt_issue_efficiency": round(c["issue_active_pct"] / 100.0 * c["lanes_per_instruction"] / 32.0, 4),
                    "what": "issue-slot utilisation x active lanes / 32 of trace_closest_kernel, instruction-weighted over the launches of one pass"}
    (REPO / "profiles" / "traversal_profile.json").write_text(json.dumps(out, indent=1) + "\n")
    print(json.dumps({k: v for k, v in out.items() if k not in ("closest", "shadow")}, indent=1))
    for kind in ("closest", "shadow"):
        print(kind, {k: v for k, v in out[kind].items() if k != "per_launch"})


if __name__ == "__main__":
    main()
