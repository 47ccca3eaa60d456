"""Aggregate an `ncu --metrics gpu__time_duration.sum[,dram__bytes_read.sum,dram__bytes_write.sum]
--csv` launch list by kernel name (markdown table on stdout).

usage: summarize_launches.py launches.csv [traffic.json [round-tag]]
With a second argument, DRAM bytes per launch of the C-ABI entry points are written as JSON
(bench.py reads profiles/r1_traffic.json for `roofline.traffic`)."""
import csv, collections, json, re, sys

UNIT_MS = {'ns': 1e-6, 'nsecond': 1e-6, 'us': 1e-3, 'usecond': 1e-3, 'ms': 1.0, 'msecond': 1.0, 's': 1e3,
           'second': 1e3}
UNIT_B = {'byte': 1.0, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9}
# kernel-name prefix -> C-ABI entry point (bench.py kernel_shares key)
ENTRY = (('tc::gcn_tc3_kernel', 'gcn_tc'), ('tc::gcn_tc2_kernel', 'gcn_tc'), ('tc::gcn_tc_kernel', 'gcn_tc'),
         ('tc::gcn_tc_dw', 'gcn_tc_dw'), ('tc::frame_colsum', 'gcn_tc_dw'), ('tc::gcn_tc_da', 'gcn_tc_dvals'),
         ('tc::gcn_pair_tc_kernel', 'gcn_pair_grads'), ('tc::pair_reduce_', 'gcn_pair_grads'),
         ('tcn_bwd', 'tcn_bwd'), ('tcn_down', 'tcn_fwd'), ('tcn_up', 'tcn_fwd'),
         ('tcn2_down_kernel', 'tcn2_down'), ('tcn2_up_kernel', 'tcn2_up'), ('tcn2_bwd_up_kernel', 'tcn2_bwd_up'),
         ('tcn2_bwd_down_kernel', 'tcn2_bwd_down'), ('tcn2_small_conv_kernel<1, 0', 'tcn2_conv'),
         ('tcn2_small_conv_kernel<2, 0', 'tcn2_conv'), ('tcn2_small_conv_kernel<1, 1', 'tcn2_bwd_conv'),
         ('tcn2_small_conv_kernel<2, 1', 'tcn2_bwd_conv'), ('tcn2_small_dw_kernel', 'tcn2_bwd_conv'),
         ('block_tail_fwd', 'block_tail_fwd'), ('block_tail_bwd', 'block_tail_bwd'),
         ('bn_back_apply', 'bn_back_apply'), ('bn_back_colsum', 'bn_back_colsum'), ('gcn_small_fwd', 'gcn_small_fwd'),
         ('gcn_small_bwd', 'gcn_small_bwd'))

with open(sys.argv[1], newline='') as f:
    lines = [l for l in f if not l.startswith('==')]
launches = collections.OrderedDict()            # id -> [name, ms, read, write]
for r in csv.DictReader(lines):
    name = re.sub(r'^void ', '', re.sub(r'\(.*', '', r['Kernel Name']))
    rec = launches.setdefault(r['ID'], [name, 0.0, 0.0, 0.0])
    v = float(r['Metric Value'].replace(',', ''))
    unit = r.get('Metric Unit', '')
    m = r.get('Metric Name')
    if m == 'gpu__time_duration.sum':
        rec[1] = v * UNIT_MS.get(unit, 1e-6)
    elif m == 'dram__bytes_read.sum':
        rec[2] = v * UNIT_B.get(unit, 1.0)
    elif m == 'dram__bytes_write.sum':
        rec[3] = v * UNIT_B.get(unit, 1.0)
rows = collections.OrderedDict()
for name, ms, rd, wr in launches.values():
    d = rows.setdefault(name, [0, 0.0, 0.0, 0.0])
    d[0] += 1; d[1] += ms; d[2] += rd; d[3] += wr
total = sum(v[1] for v in rows.values())
print('| kernel | launches | total ms | share | avg ms | DRAM read MB/launch | DRAM write MB/launch | DRAM TB/s |')
print('|---|---|---|---|---|---|---|---|')
for name, (n, t, rd, wr) in sorted(rows.items(), key=lambda kv: -kv[1][1]):
    if t / total < 0.002:
        continue
    print('| %s | %d | %.3f | %.1f%% | %.3f | %.1f | %.1f | %.2f |' % (
        name[:70], n, t, 100 * t / total, t / n, rd / n / 1e6, wr / n / 1e6, (rd + wr) / t / 1e9 if t else 0))
print('| **total (all kernels)** | %d | %.3f | | | | | |' % (sum(v[0] for v in rows.values()), total))
if len(sys.argv) > 2:
    ent = {}
    for name, (n, t, rd, wr) in rows.items():
        for prefix, key in ENTRY:
            if prefix in name:
                e = ent.setdefault(key, [0, 0.0, 0.0])
                e[0] += n; e[1] += rd + wr; e[2] += t
                break
    per_call = {'tcn_bwd': 3, 'tcn_fwd': 2, 'tcn2_bwd_conv': 2, 'gcn_pair_grads': 3}       # kernels per C-ABI call
    out = {}
    for key, (n, b, t) in ent.items():
        calls = n / per_call.get(key, 1)
        if key == 'gcn_tc_dw':                    # a frame_colsum launch may ride along: count the dw kernels
            calls = sum(v[0] for k2, v in rows.items() if k2.startswith('tc::gcn_tc_dw')) or n
        out[key] = {'dram_bytes_per_launch': b / calls, 'launches': int(calls), 'ms_per_launch': t / calls,
                    'workload': 'ntu', 'clips_per_gpu': 64,
                    'source': 'profiles/%s_launches_dram.md (ncu dram__bytes_read.sum + dram__bytes_write.sum)' % (sys.argv[3] if len(sys.argv) > 3 else 'r1')}
    with open(sys.argv[2], 'w') as f:
        json.dump(out, f, indent=1)
