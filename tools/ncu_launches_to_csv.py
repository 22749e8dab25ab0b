"""Turn an ncu launch list (`ncu --metrics gpu__time_duration.sum --csv --log-file X python bench.py --eager ...`) into the
per-step summary committed under profiles/: the launches of the LAST complete fwd+bwd step (from the last
pack_params_kernel on), one row per launch with its share of the step.
    python tools/ncu_launches_to_csv.py gpurun_out/launches.csv "header comment" > profiles/rN_step_launches.csv"""
import csv
import sys


def main():
    path = sys.argv[1]
    note = sys.argv[2] if len(sys.argv) > 2 else ""
    rows = []
    with open(path, newline="") as f:
        lines = [ln for ln in f if not ln.startswith("==")]
    rd = csv.DictReader(lines)
    for r in rd:
        if r.get("Metric Name") != "gpu__time_duration.sum":
            continue
        val = float(r["Metric Value"].replace(",", ""))
        unit = r.get("Metric Unit", "us")
        us = val / 1000.0 if unit in ("ns", "nsecond") else val * 1000.0 if unit in ("ms", "msecond") else val
        rows.append((r["Kernel Name"], r["Grid Size"], r["Block Size"], us))
    starts = [i for i, r in enumerate(rows) if "pack_params_kernel" in r[0]]
    if len(starts) >= 2:
        rows = rows[starts[-2]:starts[-1]]            # the last step that is certainly complete
    elif starts:
        rows = rows[starts[-1]:]
    tot = sum(r[3] for r in rows)
    out = csv.writer(sys.stdout, lineterminator="\n")
    for ln in note.split("\\n"):
        if ln:
            print("# " + ln)
    out.writerow(["kernel", "grid", "block", "us", "share"])
    for k, g, b, us in rows:
        out.writerow([k[:120], g, b, f"{us:.1f}", f"{us / tot:.3f}"])
    out.writerow(["total", "", "", f"{tot:.1f}", "1.000"])


if __name__ == "__main__":
    main()
