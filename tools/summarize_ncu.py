#!/usr/bin/env python
"""tools/summarize_ncu.py REPORT.ncu-rep OUT_PREFIX -- turn an `ncu --set full` report into the text evidence kept
under profiles/: key metrics per kernel (CSV), warp-stall mix from the source page, and a traffic JSON
(dram bytes per launch) that bench.py reads for roofline.traffic.  Runs on the CPU box (ncu -i)."""
import collections
import csv
import json
import subprocess
import sys

KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__grid_size",
        "launch__block_size", "launch__cluster_size", "launch__occupancy_limit_registers",
        "launch__occupancy_limit_shared_mem", "smsp__inst_executed.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct",
        "sm__cycles_elapsed.max", "smsp__inst_executed_op_local_ld.sum", "smsp__inst_executed_op_local_st.sum"]


def to_bytes(val, unit):
    v = float(val)
    return v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}.get(unit, 1)


def main():
    rep, out = sys.argv[1], sys.argv[2]
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units = rows[0], rows[1]
    ix = {h: i for i, h in enumerate(hdr)}
    traffic = {}
    with open(out + "_metrics.csv", "w") as f:
        w = csv.writer(f)
        w.writerow(["kernel"] + KEYS)
        w.writerow(["(unit)"] + [units[ix[k]] if k in ix else "" for k in KEYS])
        for r in rows[2:]:
            name = r[ix["Kernel Name"]]
            w.writerow([name] + [r[ix[k]] if k in ix else "" for k in KEYS])
            rd = to_bytes(r[ix["dram__bytes_read.sum"]], units[ix["dram__bytes_read.sum"]])
            wr = to_bytes(r[ix["dram__bytes_write.sum"]], units[ix["dram__bytes_write.sum"]])
            traffic.setdefault(name, []).append({"dram_read_bytes": rd, "dram_write_bytes": wr,
                                                 "duration_us": float(r[ix["gpu__time_duration.sum"]])})
    json.dump(traffic, open(out + "_traffic.json", "w"), indent=1)
    src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True,
                         text=True).stdout
    rows = list(csv.reader(src.splitlines()))
    hdr = rows[1]
    ix = {h: i for i, h in enumerate(hdr)}
    stall_cols = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]

    def f(x):
        try:
            return float(x)
        except ValueError:
            return 0.0
    data = [r for r in rows[2:] if len(r) == len(hdr) and r[0] != "Address"]
    seen, first = set(), []
    for r in data:  # first kernel instance only
        if r[ix["Address"]] in seen:
            break
        seen.add(r[ix["Address"]])
        first.append(r)
    tot = collections.Counter()
    for r in first:
        for c in stall_cols:
            tot[c] += f(r[ix[c]])
    s = sum(tot.values()) or 1.0
    with open(out + "_stalls.txt", "w") as fo:
        fo.write(f"# warp stall sampling, first profiled launch of {rep}\n")
        for c, v in tot.most_common():
            fo.write(f"{c:28s} {100 * v / s:5.1f} %\n")
        fo.write("\n# instructions holding >= 1.2 % of all samples (index, samples, SASS, top stall reasons)\n")
        for i, r in enumerate(first):
            n = f(r[ix["# Samples"]])
            if n >= 0.012 * s:
                st = sorted(((c, f(r[ix[c]])) for c in stall_cols), key=lambda kv: -kv[1])[:2]
                fo.write(f"{i:5d} {int(n):7d}  {r[ix['Source']][:70]:70s} {[(k[6:], int(v)) for k, v in st]}\n")
    print(open(out + "_stalls.txt").read()[:1500])


if __name__ == "__main__":
    main()
