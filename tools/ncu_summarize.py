"""Turns the CSV of `ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none
-k regex:lea_ --csv --log-file X python bench.py --batch 1 --steps 1 --warmup 3 --no-graph --quick` into the
per-launch / per-kernel summary kept under profiles/ (the LAST forward of the run: from the last feature-stem launch to
the disparity head).  Usage: python tools/ncu_summarize.py launches.csv out.json [conv_tc_traffic.json [n_matching_conv_launches]]"""
import csv
import json
import re
import sys


def short(name: str) -> str:
    name = name.replace("<unnamed>::", "").replace("void ", "")
    m = re.match(r"([A-Za-z0-9_]+)(<[^>]*>)?", name)
    base = m.group(1)
    targs = m.group(2) or ""
    targs = targs.replace("(int)", "").replace(" ", "")
    return base + targs


def main():
    src, out = sys.argv[1], sys.argv[2]
    rows = {}
    with open(src, newline="") as f:
        lines = [ln for ln in f if ln.startswith('"')]
    rd = csv.reader(lines)
    hdr = next(rd)
    ix = {h: i for i, h in enumerate(hdr)}
    for r in rd:
        i = int(r[ix["ID"]])
        e = rows.setdefault(i, {"kernel": short(r[ix["Kernel Name"]]), "grid": r[ix["Grid Size"]]})
        e[r[ix["Metric Name"]]] = float(r[ix["Metric Value"]].replace(",", ""))
    ids = sorted(rows)
    starts = [i for i in ids if rows[i]["kernel"].startswith("lea_feature_stem_kernel")]
    first = starts[-1] if starts else ids[0]
    ends = [i for i in ids if i >= first and rows[i]["kernel"].startswith("lea_disp_head")]
    last = ends[0] if ends else ids[-1]          # the forward's own head (bench.py times stand-alone kernels after it)
    sel = [rows[i] for i in ids if first <= i <= last]
    per = [{"kernel": e["kernel"], "grid": e["grid"], "us": round(e["gpu__time_duration.sum"] / 1e3, 2),
            "dram_read_MB": round(e.get("dram__bytes_read.sum", 0.0) / 1e6, 2),
            "dram_write_MB": round(e.get("dram__bytes_write.sum", 0.0) / 1e6, 2)} for e in sel]
    total = sum(p["us"] for p in per)
    by = {}
    for p in per:
        b = by.setdefault(p["kernel"], {"launches": 0, "us": 0.0, "dram_read_MB": 0.0, "dram_write_MB": 0.0})
        b["launches"] += 1
        for k in ("us", "dram_read_MB", "dram_write_MB"):
            b[k] = round(b[k] + p[k], 2)
    for b in by.values():
        b["share"] = round(b["us"] / total, 4)
    # matching-net convs: tensor-core launches after the last feature-net launch (the feature net runs 3 planes)
    stem = [k for k, p in enumerate(per) if p["kernel"].startswith("lea_conv_tc_kernel") and p["kernel"].endswith(",2>")]
    convs = [per[k] for k in stem]
    if len(sys.argv) > 4:        # number of conv launches of the matching-net plan: the LAST n tensor-core launches
        tc = [p for p in per if p["kernel"].startswith("lea_conv_tc_kernel")]
        convs = tc[-int(sys.argv[4]):]
    nbytes = sum((p["dram_read_MB"] + p["dram_write_MB"]) * 1e6 for p in convs)
    summary = {
        "command": "ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none "
                   "-k regex:lea_ --csv python bench.py --batch 1 --steps 1 --warmup 3 --no-graph --quick",
        "note": "one eager forward (1 KITTI pair) late in the run; per-launch times are cold-cache and serialised - "
                "compare shares, not absolutes",
        "launches_in_forward": len(per), "total_us": round(total, 1), "by_kernel": by,
        "matching_net_convs": {"launches": len(convs), "us": round(sum(p["us"] for p in convs), 1),
                               "share_of_forward": round(sum(p["us"] for p in convs) / total, 4),
                               "dram_bytes_per_launch_avg": int(nbytes / max(1, len(convs))), "dram_bytes_total": int(nbytes)},
        "per_launch": per}
    json.dump(summary, open(out, "w"), indent=1)
    if len(sys.argv) > 3:
        json.dump({"dram_bytes_per_launch": summary["matching_net_convs"]["dram_bytes_per_launch_avg"],
                   "launches": len(convs), "source": out}, open(sys.argv[3], "w"), indent=1)
    print(json.dumps({k: v for k, v in summary.items() if k not in ("per_launch", "by_kernel")}))


if __name__ == "__main__":
    main()
