"""Per-layer roofline table from gpurun_out/per_op_ms.json (bench.py): time vs max(FLOPs/peak, bytes/HBM)."""
import json, sys
d = json.load(open(sys.argv[1] if len(sys.argv) > 1 else 'gpurun_out/per_op_ms.json'))
peaks = json.load(open('MEASURED_PEAKS.json'))
mma_per_mac = 3.0 if 'f16x2' in (sys.argv[1] if len(sys.argv) > 1 else '') else 1.0      # the pair path issues three MMAs per algorithmic MAC
pf, pb = peaks['bf16_tflops_sustained'] * 1e12 / mma_per_mac, peaks['hbm_gbs'] * 1e9
rows, lost_total = [], 0.0
for name, ms in d['ops']:
    info = d.get('info', {}).get(name)
    if not info: continue
    t_c, t_m = info['flops'] / pf * 1e3, info['bytes'] / pb * 1e3
    bound = max(t_c, t_m)
    rows.append((ms - bound, name, ms, t_c, t_m, info['m'], info['n'], info['k']))
    lost_total += ms - bound
print('%-28s %7s %7s %7s %6s  %8s %5s %5s' % ('layer', 'ms', 'cmp_ms', 'mem_ms', 'eff', 'M', 'N', 'K'))
for lost, name, ms, t_c, t_m, m, n, k in sorted(rows, reverse=True)[:int(sys.argv[2]) if len(sys.argv) > 2 else 30]:
    print('%-28s %7.3f %7.3f %7.3f %5.0f%%  %8d %5d %5d  %s' % (name, ms, t_c, t_m, 100 * max(t_c, t_m) / ms, m, n, k, 'MEM' if t_m > t_c else 'CMP'))
print('conv total %.3f ms, roofline-bound total %.3f ms, lost %.3f ms' % (sum(r[2] for r in rows), sum(max(r[3], r[4]) for r in rows), lost_total))
