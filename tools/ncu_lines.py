"""Summarise an ncu report by CUDA source line: share of stall samples, of executed instructions, top stall reasons.

  python tools/ncu_lines.py gpurun_out/x.ncu-rep [top_n]

Reads `ncu -i <rep> --page source --print-source cuda,sass --csv` (needs -lineinfo and --import-source on).
"""
import csv
import subprocess
import sys


def f(x):
    try:
        return float(x)
    except ValueError:
        return 0.0


def main():
    rep = sys.argv[1]
    top = int(sys.argv[2]) if len(sys.argv) > 2 else 50
    txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--print-source", "cuda,sass", "--csv"],
                         capture_output=True, text=True).stdout
    rows = list(csv.reader(txt.splitlines()))
    cur, hdr, out = None, None, []
    for r in rows:
        if len(r) >= 2 and r[0] == "File Path":
            cur = r[1]
            continue
        if r and r[0] == "Line No":
            hdr = r
            continue
        if cur and hdr and len(r) > 8 and r[0].isdigit():
            d = dict(zip(hdr[4:], r[4:]))
            d["line"], d["src"], d["file"] = r[0], r[1], cur.split("/")[-1]
            out.append(d)
    stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
    ts = sum(f(d["# Samples"]) for d in out) or 1.0
    ti = sum(f(d["Instructions Executed"]) for d in out) or 1.0
    print(f"total samples {ts:.0f}  warp instructions {ti:.0f}")
    agg = {s: sum(f(d[s]) for d in out) for s in stalls}
    print("stall mix: " + ", ".join(f"{k[6:]} {100 * v / ts:.1f}%" for k, v in sorted(agg.items(), key=lambda kv: -kv[1])[:8]))
    out.sort(key=lambda d: -f(d["# Samples"]))
    for d in out[:top]:
        best = sorted(((f(d[s]), s[6:]) for s in stalls), reverse=True)[:2]
        why = " ".join(f"{n}:{100 * v / max(f(d['# Samples']), 1):.0f}%" for v, n in best)
        print(f"{d['file']:16s}{d['line']:>5s} smp {100 * f(d['# Samples']) / ts:5.1f}% ins {100 * f(d['Instructions Executed']) / ti:5.1f}% "
              f"{why:28s}| {d['src'].strip()[:100]}")


if __name__ == "__main__":
    main()
