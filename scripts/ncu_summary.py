#!/usr/bin/env python
"""Selected columns of `ncu -i X.ncu-rep --page raw --csv` -> one CSV row per captured launch + a JSON of DRAM traffic.

    python scripts/ncu_summary.py out.csv traffic.json name=rep.ncu-rep ...
"""
import csv
import io
import json
import subprocess
import sys

COLS = ['Kernel Name', 'Grid Size', 'Block Size', 'gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'dram__throughput.avg.pct_of_peak_sustained_elapsed', 'lts__t_sector_hit_rate.pct', 'lts__t_bytes.sum',
        'sm__throughput.avg.pct_of_peak_sustained_elapsed', 'gpu__compute_memory_throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active',
        'smsp__inst_issued.avg.pct_of_peak_sustained_active', 'smsp__issue_active.avg.pct', 'launch__registers_per_thread',
        'launch__shared_mem_per_block_dynamic', 'l1tex__t_sector_hit_rate.pct', 'sm__cycles_active.avg',
        'smsp__inst_executed.sum', 'lts__t_sectors_srcunit_tex_op_read.sum', 'lts__t_sectors_srcunit_tex_op_red.sum']


def to_bytes(v, unit):
    v = float(v.replace(',', ''))
    return v * {'byte': 1, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9, 'Tbyte': 1e12}.get(unit, 1)


def to_us(v, unit):
    v = float(v.replace(',', ''))
    return v * {'ns': 1e-3, 'us': 1, 'usecond': 1, 'ms': 1e3, 'msecond': 1e3, 'nsecond': 1e-3, 'second': 1e6, 's': 1e6}.get(unit, 1)


rows_out, traffic = [], {}
for arg in sys.argv[3:]:
    name, rep = arg.split('=', 1)
    txt = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
    rd = list(csv.reader(io.StringIO(txt)))
    hdr, units, data = rd[0], rd[1], rd[2:]
    idx = {h: i for i, h in enumerate(hdr)}
    for r in data:
        row = {'capture': name}
        for c in COLS:
            if c in idx:
                row[c] = r[idx[c]] + ((' ' + units[idx[c]]) if units[idx[c]] else '')
        rows_out.append(row)
        kn = r[idx['Kernel Name']]
        dr = to_bytes(r[idx['dram__bytes_read.sum']], units[idx['dram__bytes_read.sum']])
        dw = to_bytes(r[idx['dram__bytes_write.sum']], units[idx['dram__bytes_write.sum']])
        du = to_us(r[idx['gpu__time_duration.sum']], units[idx['gpu__time_duration.sum']])
        traffic.setdefault(name, []).append({'kernel': kn, 'grid': r[idx['Grid Size']], 'dram_read': dr, 'dram_write': dw,
                                             'duration_us': du, 'lts_hit_pct': r[idx['lts__t_sector_hit_rate.pct']]})
with open(sys.argv[1], 'w', newline='') as f:
    w = csv.DictWriter(f, fieldnames=['capture'] + COLS)
    w.writeheader()
    for r in rows_out:
        w.writerow(r)
json.dump(traffic, open(sys.argv[2], 'w'), indent=1)
for k, v in traffic.items():
    for x in v:
        print(k, x['kernel'][:60], 'dur %.1f us' % x['duration_us'], 'dram R %.2f MB W %.2f MB' % (x['dram_read'] / 1e6, x['dram_write'] / 1e6),
              'L2 hit', x['lts_hit_pct'])
