"""Summarise an ncu report: key metrics per kernel + top stall locations.  usage: ncu_summary.py <rep> [topN]"""
import csv, subprocess, sys, io
rep = sys.argv[1]; topn = int(sys.argv[2]) if len(sys.argv) > 2 else 14
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw))); hdr, units = rows[0], rows[1]
keys = ['gpu__time_duration.sum','dram__bytes_read.sum','dram__bytes_write.sum',
 'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed','launch__registers_per_thread','launch__grid_size',
 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed','lts__throughput.avg.pct_of_peak_sustained_elapsed',
 'l1tex__throughput.avg.pct_of_peak_sustained_elapsed','lts__t_sector_hit_rate.pct','sm__cycles_elapsed.avg',
 'sm__cycles_elapsed.avg.per_second','smsp__inst_executed.sum','l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed']
for r in rows[2:]:
    print('=====', r[hdr.index('Kernel Name')][:100])
    for k in keys:
        if k in hdr:
            i = hdr.index(k); print(f"  {k} [{units[i]}] = {r[i]}")
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
secs = []; cur = None
for ln in src.split('\n'):
    if ln.startswith('"Kernel Name"'): cur = [ln]; secs.append(cur)
    elif cur is not None: cur.append(ln)
seen = set()
for sec in secs:
    name = sec[0][:110]
    if name in seen: continue
    seen.add(name)
    rs = list(csv.reader(sec[1:])); h = rs[0]
    si, so, ie = h.index('# Samples'), h.index('Source'), h.index('Instructions Executed')
    sc = [i for i, x in enumerate(h) if x.startswith('stall_') and 'Not Issued' not in x]
    data = [r for r in rs[1:] if len(r) == len(h)]
    tot = sum(int(r[si] or 0) for r in data) or 1
    print('=====', name, 'samples', tot)
    for r in sorted(data, key=lambda r: -int(r[si] or 0))[:topn]:
        st = sorted(((int(r[i] or 0), h[i]) for i in sc), reverse=True)[:2]
        print(f"{100*int(r[si])/tot:5.1f}% ex={r[ie]:>8} {r[so][:64]:64s} {st}")
