"""Summarise an ncu report: key raw metrics + top stall lines of the source page."""
import csv, subprocess, sys, io
rep = sys.argv[1]; topn = int(sys.argv[2]) if len(sys.argv) > 2 else 22
raw = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, vals = rows[0], rows[2]
want = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'launch__registers_per_thread', 'launch__grid_size', 'smsp__inst_executed.sum',
        'lts__t_sector_hit_rate.pct', 'lts__t_bytes.sum', 'l1tex__t_bytes.sum', 'sm__cycles_elapsed.max', 'lts__t_sectors_srcunit_tex_op_read.sum',
        'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum', 'smsp__inst_executed_pipe_lsu.sum', 'sm__inst_executed_pipe_uniform.sum']
for h, u, v in zip(hdr, rows[1], vals):
    if h in want: print('%-70s %s %s' % (h, v, u))
src = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv'], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
hdr = rows[1]; data = [r for r in rows[2:] if len(r) == len(hdr)]
ix = {h: i for i, h in enumerate(hdr)}
stalls = [h for h in hdr if h.startswith('stall_') and 'Not Issued' not in h]
agg = {s: sum(int(r[ix[s]] or 0) for r in data) for s in stalls}
tot = sum(int(r[ix['# Samples']] or 0) for r in data)
print('samples', tot, sorted(agg.items(), key=lambda x: -x[1])[:7])
for r in sorted(data, key=lambda r: -int(r[ix['# Samples']] or 0))[:topn]:
    st = sorted(((s, int(r[ix[s]] or 0)) for s in stalls), key=lambda x: -x[1])[:2]
    print('%6s %-4s %-70s %s' % (r[ix['# Samples']], r[ix['Instructions Executed']][:9], r[ix['Source']][:70], st))
