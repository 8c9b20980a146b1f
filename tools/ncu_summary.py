#!/usr/bin/env python
"""Turn an `ncu --set full` report into the small JSON summaries kept under profiles/.

    python tools/ncu_summary.py gpurun_out/prof.ncu-rep profiles/r01_v1 [--match head_] [--all]

--all: one file per captured LAUNCH (suffix _launchN) instead of the first launch of each distinct kernel.

Writes one <prefix>_<short kernel name>.ncu_summary.json per distinct kernel (first captured launch of each)
with the metrics DESIGN.md quotes (DRAM bytes, durations, pipe utilisation, stall reasons, registers,
occupancy), and prints a one-line digest per kernel.  Runs here (no GPU needed): it only reads the report.
"""

import csv
import io
import json
import re
import subprocess
import sys

KEEP = re.compile(
    r'^(dram__bytes_(read|write)\.sum(\.per_second|\.pct_of_peak_sustained_elapsed)?|gpu__time_duration\.sum|'
    r'gpu__dram_throughput\.avg\.pct_of_peak_sustained_elapsed|launch__(block_size|grid_size|registers_per_thread|'
    r'occupancy_limit_\w+|waves_per_multiprocessor)|sm__cycles_elapsed\.avg\.per_second|'
    r'sm__inst_executed_pipe_(alu|fma|fmaheavy|lsu|xu|uniform)\.avg\.pct_of_peak_sustained_active|'
    r'sm__throughput\.avg\.pct_of_peak_sustained_elapsed|sm__warps_active\.avg\.pct_of_peak_sustained_active|'
    r'smsp__average_warps_issue_stalled_\w+_per_issue_active\.ratio|smsp__inst_executed\.sum|'
    r'smsp__issue_active\.avg\.pct_of_peak_sustained_active|lts__t_bytes\.sum|lts__t_sector_hit_rate\.pct|'
    r'l1tex__t_bytes_pipe_lsu_mem_global_op_(ld|st)\.sum|lts__t_sectors_srcunit_tex_op_(read|write)\.sum|'
    r'sm__maximum_warps_per_active_cycle_pct|lts__throughput\.avg\.pct_of_peak_sustained_elapsed)$')


def short_name(full):
    m = re.search(r'(\w+)<([^>]*)>', full)
    if not m:
        return re.sub(r'\W+', '_', full)[:60]
    args = re.sub(r'[^0-9A-Za-z]+', '_', m.group(2).replace('__nv_bfloat16', 'bf16').replace('float', 'f32')).strip('_')
    return '%s_%s' % (m.group(1), args)


def to_number(text):
    try:
        return float(text.replace(',', ''))
    except ValueError:
        return text


def main():
    if len(sys.argv) < 3:
        print(__doc__)
        return 2
    rep, prefix = sys.argv[1], sys.argv[2]
    match = sys.argv[sys.argv.index('--match') + 1] if '--match' in sys.argv else ''
    every = '--all' in sys.argv
    raw = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], check=True, capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    header, units, body = rows[0], rows[1], rows[2:]
    name_col = header.index('Kernel Name')
    seen = {}
    for idx, row in enumerate(body):
        kname = row[name_col]
        if match and match not in kname:
            continue
        if kname in seen and not every:
            continue
        summary = {'Kernel Name': kname}
        for col, unit, val in zip(header, units, row):
            if KEEP.match(col):
                summary[col] = {'value': to_number(val), 'unit': unit}
        seen[kname] = summary
        out = '%s_%s%s.ncu_summary.json' % (prefix, short_name(kname), '_launch%d' % idx if every else '')
        with open(out, 'w') as f:
            json.dump(summary, f, indent=1, sort_keys=True)

        def g(key):
            v = summary.get(key, {}).get('value')
            return v if isinstance(v, float) else float('nan')
        traffic = g('dram__bytes_read.sum') + g('dram__bytes_write.sum')
        print('%s\n   -> %s\n   duration %.1f %s | dram r+w %.3f %s | dram %.1f%% | issue-active %.1f%% | regs %d | warps-active %.1f%%'
              % (kname, out, g('gpu__time_duration.sum'), summary.get('gpu__time_duration.sum', {}).get('unit'),
                 traffic, summary.get('dram__bytes_read.sum', {}).get('unit'),
                 g('gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed'),
                 g('smsp__issue_active.avg.pct_of_peak_sustained_active'), int(g('launch__registers_per_thread')),
                 g('sm__warps_active.avg.pct_of_peak_sustained_active')))
    return 0


if __name__ == '__main__':
    sys.exit(main())
