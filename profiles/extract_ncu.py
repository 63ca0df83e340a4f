"""Condenses an `ncu --set full` report into one markdown table (one row per captured launch):
    python profiles/extract_ncu.py gpurun_out/x.ncu-rep "title" > profiles/x.md
Reads the report with `ncu -i ... --page raw --csv` (works without a GPU)."""
import csv
import io
import subprocess
import sys

COLS = [("gpu__time_duration.sum", "us", 1e-3), ("dram__bytes_read.sum", "dram rd MB", 1e-6),
        ("dram__bytes_write.sum", "dram wr MB", 1e-6), ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram %", 1),
        ("lts__t_sector_hit_rate.pct", "L2 hit %", 1), ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "L2 %", 1),
        ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor % (active)", 1),
        ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "SM %", 1),
        ("sm__warps_active.avg.pct_of_peak_sustained_active", "occupancy %", 1),
        ("launch__registers_per_thread", "regs", 1), ("launch__grid_size", "grid", 1)]


def main():
    rep, title = sys.argv[1], sys.argv[2] if len(sys.argv) > 2 else sys.argv[1]
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], stdout=subprocess.PIPE, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    head = rows[0]
    units = rows[1]
    idx = {h: i for i, h in enumerate(head)}
    name_i = idx.get("Kernel Name")
    print("### %s\n" % title)
    print("| kernel | " + " | ".join(c[1] for c in COLS) + " |")
    print("|---|" + "---:|" * len(COLS))
    for r in rows[2:]:
        if len(r) <= name_i:
            continue
        cells = []
        for m, _, scale in COLS:
            i = idx.get(m)
            if i is None or r[i] in ("", "n/a"):
                cells.append("-")
                continue
            v = float(r[i].replace(",", ""))
            u = units[i]
            if m == "gpu__time_duration.sum":
                v = v * {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}.get(u, 1e-3)
            elif m.startswith("dram__bytes"):
                v = v * {"byte": 1e-6, "Kbyte": 1e-3, "Mbyte": 1.0, "Gbyte": 1e3}.get(u, 1e-6)
            cells.append("%.1f" % v if abs(v) < 1e5 else "%.0f" % v)
        name = r[name_i].split("(")[0].replace("fv2p::<unnamed>::", "")
        print("| `%s` | " % name[:60] + " | ".join(cells) + " |")


if __name__ == "__main__":
    main()
