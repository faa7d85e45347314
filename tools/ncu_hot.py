"""Summarise an ncu report: key raw metrics + per-phase / per-instruction stall samples (source page)."""
import csv, subprocess, sys, io
rep = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 30
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
want = ['Kernel Name', 'gpu__time_duration.sum', 'launch__registers_per_thread', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'lts__throughput.avg.pct_of_peak_sustained_elapsed',
        'l1tex__throughput.avg.pct_of_peak_sustained_elapsed', 'smsp__inst_executed.sum', 'sass__inst_executed_local_loads',
        'sass__inst_executed_local_stores', 'lts__t_sectors_srcunit_tex_op_red.sum', 'launch__shared_mem_per_block_dynamic',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'launch__grid_size', 'launch__block_size',
        'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum', 'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum']
for vals in rows[2:]:
    for h, u, v in zip(hdr, units, vals):
        if h in want: print(h, u, v)
        elif 'pcsamp_warps_issue_stalled' in h and 'not_issued' not in h:
            try:
                if float(v) >= 3000: print(h, v)
            except ValueError: pass
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
hdr = rows[1]; data = rows[2:]
isrc = hdr.index('Source'); ins = hdr.index('# Samples'); iex = hdr.index('Instructions Executed')
il = hdr.index('stall_long_sb'); iw = hdr.index('stall_wait'); ish = hdr.index('stall_short_sb')
f = lambda r, i: int(r[i] or 0)
print('total samples', sum(f(r, ins) for r in data), 'instructions', len(data))
bars = [i for i, r in enumerate(data) if 'BAR.' in r[isrc]]
prev = 0
for b in bars + [len(data)]:
    seg = data[prev:b]
    print(f'[{prev},{b}) samples', sum(f(r, ins) for r in seg), 'warp-instr', sum(f(r, iex) for r in seg),
          'long', sum(f(r, il) for r in seg), 'wait', sum(f(r, iw) for r in seg), 'short', sum(f(r, ish) for r in seg))
    prev = b
idx = sorted(range(len(data)), key=lambda i: -f(data[i], ins))[:top]
for i in sorted(idx):
    r = data[i]
    print(i, r[ins], 'L', r[il], 'W', r[iw], 'S', r[ish], r[iex], r[isrc][:90])
