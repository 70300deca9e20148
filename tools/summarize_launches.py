"""ncu launch-list CSV (one eager step) -> compact per-launch table with plan-step names and kernel shares.
Optional third argument: path of a JSON summary of the conv kernel family (DRAM bytes and time share of one step) that bench.py
reports as `roofline.traffic`."""
import csv, json, sys, collections
path, names_path = sys.argv[1], sys.argv[2]
rows = list(csv.reader(open(path)))
hi = [i for i, r in enumerate(rows) if r and r[0] == 'ID'][0]
hdr, data = rows[hi], rows[hi + 1:]
ix = {h: i for i, h in enumerate(hdr)}
per = collections.OrderedDict()
for r in data:
    if len(r) < len(hdr): continue
    per.setdefault(int(r[ix['ID']]), {'name': r[ix['Kernel Name']]})[r[ix['Metric Name']]] = float(r[ix['Metric Value']].replace(',', ''))
names = open(names_path).read().split('\n')
# dense post-processing: every decode step is two launches (anchor math, then the streaming score kernel with the fused
# score histogram), matrix_nms is cutoff + collect + matrix (its memset is not a kernel)
plan = []
for n in names:
    if n == 'matrix_nms': plan += [n + ':cutoff', n + ':collect', n + ':matrix']
    elif n.startswith('decode'): plan += [n + ':anchors', n + ':scores']
    else: plan.append(n)
tot = sum(m['gpu__time_duration.sum'] for m in per.values())
print('| # | plan step | kernel | time us | share | DRAM rd MB | DRAM wr MB | tensor pipe % | warps active % | grid | regs |')
print('|---|---|---|---|---|---|---|---|---|---|---|')
fam = collections.Counter()
for k, m in per.items():
    kn = m['name'].split('(')[0].replace('void ', '').replace('unnamed>::', '').replace('<unnamed>::', '')
    fam[kn.split('<')[0]] += m['gpu__time_duration.sum']
    print('| %d | %s | %s | %.1f | %.1f%% | %.1f | %.1f | %.1f | %.1f | %d | %d |' % (
        k, plan[k] if k < len(plan) else '?', kn[:44], m['gpu__time_duration.sum'] / 1e3, 100 * m['gpu__time_duration.sum'] / tot,
        m['dram__bytes_read.sum'] / 1e6, m['dram__bytes_write.sum'] / 1e6,
        m['sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active'], m['sm__warps_active.avg.pct_of_peak_sustained_active'],
        m['launch__grid_size'], m['launch__registers_per_thread']))
print()
print('total %.1f us (cold-cache, serialised ncu replay: compare shares, not absolutes)' % (tot / 1e3))
for f, t in fam.most_common():
    print('  %-28s %.1f us  %.1f%%' % (f, t / 1e3, 100 * t / tot))

if len(sys.argv) > 3:
    conv = [m for m in per.values() if 'conv_umma_kernel' in m['name'] or 'dcn_umma_kernel' in m['name']]
    t_conv = sum(m['gpu__time_duration.sum'] for m in conv)
    json.dump({'kernel_family': 'conv_umma_kernel + dcn_umma_kernel launches of one eager step (ncu, --clock-control none)',
               'launches': len(conv), 'dram_bytes_read': sum(m['dram__bytes_read.sum'] for m in conv),
               'dram_bytes_write': sum(m['dram__bytes_write.sum'] for m in conv),
               'dram_bytes': sum(m['dram__bytes_read.sum'] + m['dram__bytes_write.sum'] for m in conv),
               'time_share_of_step': t_conv / tot, 'source_csv': path}, open(sys.argv[3], 'w'), indent=1)
