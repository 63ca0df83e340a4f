"""Extracts the headline metrics of one `ncu --set full` capture (first kernel in the report) to a text file and
prints dram bytes per launch.   python profiles/extract_ncu.py <report.ncu-rep> <out.txt> "<title>" """
import csv
import io
import subprocess
import sys

rep, out_path, title = sys.argv[1], sys.argv[2], sys.argv[3]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units, vals = rows[0], rows[1], rows[2]
WANT = ["dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "gpu__time_duration.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "launch__block_size",
        "launch__grid_size", "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic",
        "lts__t_sector_hit_rate.pct", "lts__t_sectors_srcunit_tex.sum",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed", "sm__cycles_active.avg",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__inst_executed.sum"]
lines = ["# ncu --set full --clock-control none: " + title, "# kernel: " + vals[hdr.index("Kernel Name")][:110]]
for w in WANT:
    if w in hdr:
        i = hdr.index(w)
        lines.append("%-75s %-16s %s" % (w, units[i], vals[i]))
open(out_path, "w").write("\n".join(lines) + "\n")


def to_bytes(name):
    i = hdr.index(name)
    v = float(vals[i].replace(",", ""))
    return v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[units[i]]


print(int(to_bytes("dram__bytes_read.sum") + to_bytes("dram__bytes_write.sum")))
