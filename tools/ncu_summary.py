"""Summarise an ncu report (one or more kernels, --set full; .ncu-rep or its `--page raw --csv` export) into a table:
python tools/ncu_summary.py rep.ncu-rep|raw.csv [peak_GBs]"""
import csv
import subprocess
import sys

rep = sys.argv[1]
peak = float(sys.argv[2]) if len(sys.argv) > 2 else 6546.2
txt = open(rep).read() if rep.endswith(".csv") else subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(txt.splitlines()))
hdr, units = rows[0], rows[1]
col = {h: i for i, h in enumerate(hdr)}
want = [("gpu__time_duration.sum", "ms"), ("dram__bytes_read.sum", "read"), ("dram__bytes_write.sum", "write"),
        ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram%"),
        ("sm__pipe_tensor_subpipe_dmma_cycles_active.avg.pct_of_peak_sustained_active", "dmma%"),
        ("sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_elapsed", "fp64%"),
        ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue%"),
        ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps%"),
        ("launch__registers_per_thread", "regs"), ("launch__shared_mem_per_block_dynamic", "dsmem"),
        ("launch__grid_size", "grid"), ("launch__block_size", "block")]


def val(r, name):
    if name not in col:
        return float("nan")
    v, u = r[col[name]].replace(",", ""), units[col[name]]
    try:
        x = float(v)
    except ValueError:
        return float("nan")
    u = u.split("/")[0]
    scale = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0, "ms": 1.0, "us": 1e-3, "ns": 1e-6, "s": 1e3,
             "usecond": 1e-3, "msecond": 1.0, "nsecond": 1e-6, "second": 1e3}.get(u, 1.0)
    return x * scale


print(f"# {rep}: one row per captured launch; HBM GB/s = measured DRAM bytes / duration, frac of {peak} GB/s (MEASURED_PEAKS.json)")
print(f"{'kernel':34s} {'ms':>8s} {'DRAM MB':>9s} {'GB/s':>7s} {'frac':>6s} {'dram%':>6s} {'dmma%':>6s} {'fp64%':>6s} {'issue%':>6s} {'warps%':>6s} {'regs':>5s} {'dsmem KB':>8s} {'grid x block':>12s}")
for r in rows[2:]:
    if len(r) < len(hdr):
        continue
    name = r[col["Kernel Name"]].split("(")[0][:34]
    ms = val(r, "gpu__time_duration.sum")
    by = val(r, "dram__bytes_read.sum") + val(r, "dram__bytes_write.sum")
    gbs = by / (ms * 1e-3) / 1e9 if ms > 0 else float("nan")
    print(f"{name:34s} {ms:8.3f} {by / 1e6:9.1f} {gbs:7.0f} {gbs / peak:6.3f} {val(r, want[3][0]):6.1f} {val(r, want[4][0]):6.1f} "
          f"{val(r, want[5][0]):6.1f} {val(r, want[6][0]):6.1f} {val(r, want[7][0]):6.1f} {val(r, want[8][0]):5.0f} "
          f"{val(r, want[9][0]) / 1e3:8.1f} {int(val(r, want[10][0])):5d} x {int(val(r, want[11][0])):4d}")
