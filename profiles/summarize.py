#!/usr/bin/env python
"""Turn ncu outputs (gpurun_out/) into the small text summaries committed under profiles/.

  python profiles/summarize.py launches gpurun_out/launches.csv        > profiles/rNN_launches.txt
  python profiles/summarize.py raw      gpurun_out/prof.ncu-rep        > profiles/rNN_kernel.txt
  python profiles/summarize.py counters gpurun_out/prof.ncu-rep profiles/rNN_kernel.txt > profiles/ncu_counters.json
"""
import collections
import csv
import subprocess
import sys

KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum ", "dram__bytes_write.sum ", "dram__bytes_read.sum.per_second",
        "dram__bytes_write.sum.per_second", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread ", "launch__grid_size", "launch__block_size", "launch__occupancy_limit_registers",
        "launch__occupancy_limit_shared_mem", "launch__shared_mem_per_block_dynamic",
        "sm__inst_executed.sum.per_cycle_elapsed", "smsp__inst_issued.sum ", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
        "sm__pipe_tensor_cycles_active", "sm__inst_executed_pipe_tensor", "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum ", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct", "smsp__average_warps_issue_stalled",
        "smsp__average_warp_latency_per_inst_issued"]


def launches(path):
    lines = [l for l in open(path) if not l.startswith("==")]
    agg = collections.OrderedDict()
    for row in csv.DictReader(lines):
        if row.get("Metric Name") != "gpu__time_duration.sum":
            continue
        v = float(row["Metric Value"].replace(",", ""))
        v *= {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}.get(row["Metric Unit"], 1.0)
        a = agg.setdefault(row["Kernel Name"][:70], [0, 0.0])
        a[0] += 1
        a[1] += v
    tot = sum(a[1] for a in agg.values())
    print(f"# {path}: per-kernel device time under ncu (cold cache, serialised) -- compare SHARES")
    for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"{k:70s} n={n:4d} total={t / 1e3:10.3f} ms  avg={t / n:10.1f} us  share={t / tot:6.3f}")


def raw(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units, data = rows[0], rows[1], rows[2:]
    name_col = hdr.index("Kernel Name")
    print(f"# {path}: selected counters per captured launch (ncu --set full --clock-control none)")
    for d in data:
        print(f"\n== {d[name_col][:90]}")
        for i, h in enumerate(hdr):
            if any(h.startswith(k.strip()) if k.endswith(" ") else (k in h) for k in KEYS) and d[i] not in ("", "0"):
                if k_ok(h):
                    print(f"  {h:95s} {d[i]:>18s} {units[i]}")


def counters(path, source):
    """JSON for profiles/ncu_counters.json: per kernel instance (template arguments kept) the figures bench.py
    prints beside its own timings -- DRAM bytes per launch and the pipe / occupancy percentages."""
    import json
    import re
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units, data = rows[0], rows[1], rows[2:]
    col = {h: i for i, h in enumerate(hdr)}
    scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1e-3, "us": 1.0, "ms": 1e3, "usecond": 1.0,
             "nsecond": 1e-3, "msecond": 1e3}

    def val(d, name):
        i = col.get(name)
        if i is None or d[i] == "":
            return None
        return float(d[i].replace(",", "")) * scale.get(units[i], 1.0)

    res = {}
    for d in data:
        name = d[col["Kernel Name"]]
        m = re.search(r"(\w+)<([^>]*)>", name)
        key = f"{m.group(1)}<{m.group(2).replace(' ', '')}>" if m else re.sub(r"\(.*", "", name).split("::")[-1]
        rd, wr = val(d, "dram__bytes_read.sum"), val(d, "dram__bytes_write.sum")
        res[key] = {
            "dram_bytes": None if rd is None else rd + wr, "dram_read_bytes": rd, "dram_write_bytes": wr,
            "duration_us": val(d, "gpu__time_duration.sum"),
            "warps_active_pct": val(d, "sm__warps_active.avg.pct_of_peak_sustained_active"),
            "l1_data_pipe_pct": val(d, "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed"),
            "dmma_pipe_pct": val(d, "sm__inst_executed_pipe_tensor_subpipe_dmma.avg.pct_of_peak_sustained_active"),
            "fp64_pipe_pct": val(d, "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active"),
            "dram_pct": val(d, "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed"),
            "registers_per_thread": val(d, "launch__registers_per_thread"),
            "grid": val(d, "launch__grid_size"), "source": source}
    print(json.dumps(res, indent=1))


def k_ok(h):
    return not any(x in h for x in (".max", ".min", "peak_sustained ", ".per_second Ghz", "_not_issued"))


if __name__ == "__main__":
    if sys.argv[1] == "counters":
        counters(sys.argv[2], sys.argv[3] if len(sys.argv) > 3 else sys.argv[2])
    else:
        {"launches": launches, "raw": raw}[sys.argv[1]](sys.argv[2])
