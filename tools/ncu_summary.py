"""Summarise `ncu --page raw --csv` exports (tools/profile_r2.sh) as a markdown table: one row per captured launch with duration,
DRAM bytes, achieved HBM rate vs the measured peak, tensor-pipe / issue activity, registers.
    python tools/ncu_summary.py gpurun_out/r2_prof_fwd.csv [...] > profiles/r2_ncu_summary.md"""
import csv
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def col(header, *needles):
    for i, h in enumerate(header):
        if all(n in h for n in needles):
            return i
    return None


def to_float(s):
    try:
        return float(s.replace(",", ""))
    except (ValueError, AttributeError):
        return None


def scale(val, unit, kind):
    """normalise to ns / bytes"""
    if val is None:
        return None
    u = (unit or "").lower()
    if kind == "time":
        return val * {"ns": 1, "us": 1e3, "usecond": 1e3, "ms": 1e6, "msecond": 1e6, "s": 1e9, "second": 1e9, "nsecond": 1}.get(u, 1)
    return val * {"byte": 1, "kbyte": 1e3, "mbyte": 1e6, "gbyte": 1e9, "tbyte": 1e12}.get(u, 1)


def main():
    peak = 6573.8
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        peak = json.load(open(p))["hbm_gbs"]
    print("| file | kernel | time (ms) | DRAM read (GB) | DRAM write (GB) | DRAM GB/s | of measured HBM | tensor pipe active % | issue active % | regs | grid x block |")
    print("|---|---|---|---|---|---|---|---|---|---|---|")
    for path in sys.argv[1:]:
        rows = [r for r in csv.reader(open(path)) if r]
        start = next((i for i, r in enumerate(rows) if "Kernel Name" in r), None)
        if start is None:
            print(f"| {os.path.basename(path)} | (no kernel rows) | | | | | | | | | |")
            continue
        header, units = rows[start], rows[start + 1]
        exact = lambda n: (header.index(n) if n in header else None)
        ix = dict(name=exact("Kernel Name"), t=exact("gpu__time_duration.sum"), rd=exact("dram__bytes_read.sum"),
                  wr=exact("dram__bytes_write.sum"), tens=exact("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active"),
                  tens2=None, issue=exact("smsp__issue_active.avg.pct_of_peak_sustained_active"),
                  regs=exact("launch__registers_per_thread"), grid=exact("Grid Size"), block=exact("Block Size"))
        keep = [c for c in ("ID", "Kernel Name", "Grid Size", "Block Size", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
                            "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
                            "sm__throughput.avg.pct_of_peak_sustained_elapsed", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
                            "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "smsp__inst_executed.sum",
                            "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct",
                            "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active") if c in header]
        if os.environ.get("NCU_TRIM_DIR"):
            with open(os.path.join(os.environ["NCU_TRIM_DIR"], os.path.basename(path).replace(".csv", "_trimmed.csv")), "w", newline="") as fo:
                w = csv.writer(fo)
                for r in rows[start:]:
                    w.writerow([r[header.index(c)] if header.index(c) < len(r) else "" for c in keep])
        for r in rows[start + 2:]:
            g = lambda k: (r[ix[k]] if ix[k] is not None and ix[k] < len(r) else None)
            u = lambda k: (units[ix[k]] if ix[k] is not None else None)
            t = scale(to_float(g("t")), u("t"), "time")
            rd = scale(to_float(g("rd")), u("rd"), "bytes")
            wr = scale(to_float(g("wr")), u("wr"), "bytes")
            name = (g("name") or "")
            name = name[:name.index("(")] if "(" in name else name
            tens = to_float(g("tens")) if ix["tens"] is not None else to_float(g("tens2"))
            gbs = (rd + wr) / t if (t and rd is not None and wr is not None) else None
            f = lambda x, n=3: "" if x is None else f"{x:.{n}f}"
            print(f"| {os.path.basename(path)} | `{name}` | {f(t / 1e6 if t else None)} | {f(rd / 1e9 if rd is not None else None)} | "
                  f"{f(wr / 1e9 if wr is not None else None)} | {f(gbs, 0)} | {f(gbs / peak if gbs else None, 2)} | {f(tens, 1)} | "
                  f"{f(to_float(g('issue')), 1)} | {g('regs') or ''} | {g('grid') or ''} x {g('block') or ''} |")


if __name__ == "__main__":
    main()
