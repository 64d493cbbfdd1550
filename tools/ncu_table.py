"""One line per launch of an .ncu-rep: grid, registers, warps active, DRAM throughput, time, DRAM bytes, TB/s.
usage: ncu_table.py <rep>   (no GPU needed)"""
import csv
import io
import subprocess
import sys

raw = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
h = rows[0]


def g(r, k):
    return r[h.index(k)] if k in h else ""


print(f"{'kernel':34s} {'grid':>6s} {'regs':>4s} {'warps%':>6s} {'dram%':>6s} {'ms':>7s} {'GB':>6s} {'TB/s':>5s} waves")
for r in rows[2:]:
    gb = float(g(r, "dram__bytes_read.sum") or 0) + float(g(r, "dram__bytes_write.sum") or 0)
    ms = float(g(r, "gpu__time_duration.sum") or 0)
    print(f"{g(r, 'Kernel Name')[:34]:34s} {g(r, 'launch__grid_size'):>6s} {g(r, 'launch__registers_per_thread'):>4s} "
          f"{g(r, 'sm__warps_active.avg.pct_of_peak_sustained_active')[:5]:>6s} "
          f"{g(r, 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed')[:5]:>6s} {ms:7.3f} {gb:6.2f} "
          f"{gb / ms if ms else 0:5.2f} {g(r, 'launch__waves_per_multiprocessor')}")
print("# time unit:", rows[1][h.index("gpu__time_duration.sum")], " bytes unit:", rows[1][h.index("dram__bytes_read.sum")])
