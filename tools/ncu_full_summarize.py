"""Turns `ncu -i X.ncu-rep --page raw --csv` (a `--set full` capture of tools/ncu_target.py) into the small JSON kept
under profiles/: per profiled launch the duration, DRAM bytes, DRAM / tensor-pipe / SM / L1 utilisation, registers,
instruction count.  Usage: python tools/ncu_full_summarize.py raw.csv out.json "<command line that was captured>" """
import csv
import json
import sys

KEEP = ["gpu__time_duration.sum", "sm__cycles_elapsed.max", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_tensor.sum", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "smsp__inst_executed.sum",
        "smsp__cycles_active.avg", "sm__pipe_xu_cycles_active.avg.pct_of_peak_sustained_active"]


def main():
    src, out, cmd = sys.argv[1], sys.argv[2], (sys.argv[3] if len(sys.argv) > 3 else "")
    with open(src, newline="") as f:
        rows = list(csv.reader(ln for ln in f if ln.startswith('"')))
    hdr, units, body = rows[0], rows[1], rows[2:]
    ix = {h: i for i, h in enumerate(hdr)}
    kernels = []
    for r in body:
        e = {"launch": int(r[ix["ID"]]), "kernel": r[ix["Kernel Name"]], "grid": r[ix["Grid Size"]]}
        for k in KEEP:
            if k in ix and r[ix[k]] != "":
                e[k] = "%s %s" % (r[ix[k]], units[ix[k]])
        kernels.append(e)
    json.dump({"command": cmd, "kernels": kernels}, open(out, "w"), indent=1)
    for e in kernels:
        print(e["launch"], e["kernel"][:70], e.get("gpu__time_duration.sum"), e.get("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active"),
              e.get("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed"))


if __name__ == "__main__":
    main()
