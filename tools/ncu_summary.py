"""Turn ncu captures into the small text/JSON summaries kept under profiles/.

  python tools/ncu_summary.py full   gpurun_out/prof.ncu-rep  profiles/rN_ncu_full_summary.json
  python tools/ncu_summary.py list   gpurun_out/launches.csv  profiles/rN_launches_summary.txt

`full`  reads a `ncu --set full` report through `ncu -i REP --page raw --csv` and keeps, per profiled kernel,
        duration, DRAM bytes, occupancy, pipe utilisation, shared-memory wavefronts / bank conflicts and the
        warp-stall breakdown (pc sampling).
`list`  reads the CSV log of `ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum`
        and prints each kernel's share of the profiled launches plus the first launches with their DRAM bytes.
"""
import csv
import io
import json
import subprocess
import sys

KEEP = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
    "launch__grid_size", "launch__block_size", "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic",
    "launch__occupancy_limit_shared_mem", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "smsp__issue_active.avg.per_cycle_active", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "smsp__inst_executed.sum", "lts__t_sector_hit_rate.pct",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
]


def short(name):
    """b200ctc::<unnamed>::lattice_kernel<2, 0, 4>(b200ctc::LatticeParams) -> lattice_kernel<2, 0, 4>"""
    depth, cut = 0, len(name)
    for i, ch in enumerate(name):
        if ch == "<":
            depth += 1
        elif ch == ">":
            depth = max(0, depth - 1)
        elif ch == "(" and depth == 0:
            cut = i
            break
    name = name[:cut]
    depth, start = 0, 0
    for i, ch in enumerate(name):
        if ch == "<":
            depth += 1
        elif ch == ">":
            depth = max(0, depth - 1)
        elif ch == ":" and depth == 0:
            start = i + 1
    return name[start:][:70]


def full(rep, out):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], check=True, capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units, data = rows[0], rows[1], rows[2:]
    col = {h: i for i, h in enumerate(hdr)}
    result = []
    for r in data:
        d = {"kernel": short(r[col["Kernel Name"]]), "grid": r[col["Grid Size"]], "block": r[col["Block Size"]]}
        for k in KEEP:
            if k in col:
                d[k] = ("%s %s" % (r[col[k]], units[col[k]])).strip()
        stalls = {}
        for h, i in col.items():
            if h.startswith("smsp__pcsamp_warps_issue_stalled_") and not h.endswith("_not_issued"):
                try:
                    stalls[h[len("smsp__pcsamp_warps_issue_stalled_"):]] = float(r[i].replace(",", ""))
                except ValueError:
                    pass
        tot = sum(stalls.values()) or 1.0
        d["warp_stall_breakdown_pct"] = {k: round(100.0 * v / tot, 1)
                                         for k, v in sorted(stalls.items(), key=lambda kv: -kv[1]) if v / tot >= 0.01}
        result.append(d)
    with open(out, "w") as f:
        json.dump(result, f, indent=1)
    for d in result:
        print(d["kernel"], d.get("gpu__time_duration.sum"), d.get("dram__bytes_read.sum"), d.get("dram__bytes_write.sum"))


def launch_list(path, out):
    lines = [l for l in open(path) if l.startswith('"')]
    rows = list(csv.DictReader(io.StringIO("".join(lines))))
    per = {}
    order = []
    for r in rows:
        key = (r["ID"], short(r["Kernel Name"]))
        if key not in per:
            per[key] = {}
            order.append(key)
        v = float(r["Metric Value"].replace(",", ""))
        unit = r["Metric Unit"]
        name = r["Metric Name"]
        if name == "gpu__time_duration.sum":
            v *= {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}.get(unit, 1.0)          # -> us
        else:
            v *= {"byte": 1e-6, "Kbyte": 1e-3, "Mbyte": 1.0, "Gbyte": 1e3}.get(unit, 1.0)   # -> MB
        per[key][name] = v
    tot = {}
    for key in order:
        tot[key[1]] = tot.get(key[1], 0.0) + per[key].get("gpu__time_duration.sum", 0.0)
    total = sum(tot.values()) or 1.0
    with open(out, "w") as f:
        f.write("# per-launch times are cold-cache and serialised under ncu: compare SHARES, not absolutes\n\n")
        for k, v in sorted(tot.items(), key=lambda kv: -kv[1]):
            f.write("%-70s %9.1f us total  %5.1f %%\n" % (k, v, 100.0 * v / total))
        f.write("\n")
        for key in order[:16]:
            m = per[key]
            f.write("%-70s %8.1f us  dram read %8.1f MB  write %8.1f MB\n" % (
                key[1], m.get("gpu__time_duration.sum", 0.0), m.get("dram__bytes_read.sum", 0.0),
                m.get("dram__bytes_write.sum", 0.0)))
    print(open(out).read())


if __name__ == "__main__":
    {"full": full, "list": launch_list}[sys.argv[1]](sys.argv[2], sys.argv[3])
