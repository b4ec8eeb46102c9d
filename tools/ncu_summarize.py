'''Turns ncu output into the summaries committed under profiles/.

  launches <csv> <out.md> [title]     launch list of `ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file <csv> ...`
  kernel <file.ncu-rep> <out.json> <config text>   one `ncu --set full` capture (read with `ncu -i ... --page raw --csv`)
'''
import csv
import io
import json
import subprocess
import sys
from collections import OrderedDict


def launches(path, out, title):
    rows = []
    with open(path, newline='') as f:
        lines = [l for l in f if not l.startswith('==')]
    rd = csv.DictReader(io.StringIO(''.join(lines)))
    for r in rd:
        if r.get('Metric Name') != 'gpu__time_duration.sum':
            continue
        v = float(r['Metric Value'].replace(',', ''))
        unit = r['Metric Unit']
        ms = v * {'ns': 1e-6, 'us': 1e-3, 'ms': 1.0, 's': 1e3}.get(unit, 1e-6)
        rows.append((r['Kernel Name'], ms))
    agg = OrderedDict()
    for k, ms in rows:
        a = agg.setdefault(k, [0, 0.0])
        a[0] += 1
        a[1] += ms
    tot = sum(ms for _, ms in rows)
    with open(out, 'w') as f:
        f.write(f'# {title}\n# ncu --metrics gpu__time_duration.sum --clock-control none -c 400 ; per-launch times are cold-cache and serialised: compare SHARES\n\n')
        f.write(f'total kernel time in capture: {tot:.3f} ms over {len(rows)} launches\n\n| kernel | launches | total ms | share | avg us |\n|---|---|---|---|---|\n')
        for k, (n, ms) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            f.write(f'| `{k[:90]}` | {n} | {ms:.3f} | {100 * ms / tot:.1f}% | {1e3 * ms / n:.1f} |\n')


WANT = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'launch__registers_per_thread', 'launch__grid_size',
        'launch__block_size', 'sm__warps_active.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active',
        'smsp__issue_active.avg.pct_of_peak_sustained_active', 'sm__throughput.avg.pct_of_peak_sustained_elapsed',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'smsp__inst_executed.sum', 'launch__occupancy_limit_registers',
        'launch__occupancy_limit_shared_mem', 'smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio', 'smsp__average_warps_issue_stalled_wait_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio', 'smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio',
        'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum', 'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum', 'sm__cycles_elapsed.max',
        'sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active', 'Kernel Name']


def to_bytes(v, unit):
    return float(v.replace(',', '')) * {'byte': 1, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9}[unit]


def kernel(rep, out, config, algorithmic_bytes=None):
    txt = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True, check=True).stdout
    rd = list(csv.reader(io.StringIO(txt)))
    hdr, units, vals = rd[0], rd[1], rd[2]
    m = {h: {'value': v, 'unit': u} for h, u, v in zip(hdr, units, vals) if h in WANT}
    rdb = to_bytes(m['dram__bytes_read.sum']['value'], m['dram__bytes_read.sum']['unit'])
    wrb = to_bytes(m['dram__bytes_write.sum']['value'], m['dram__bytes_write.sum']['unit'])
    dur = float(m['gpu__time_duration.sum']['value'].replace(',', ''))
    dur_ms = dur * {'ns': 1e-6, 'us': 1e-3, 'ms': 1.0, 's': 1e3}.get(m['gpu__time_duration.sum']['unit'], 1.0)
    res = {'round': 1, 'kernel': m['Kernel Name']['value'], 'config': config, 'duration_ms_under_ncu': dur_ms,
           'dram_bytes_per_launch': rdb + wrb, 'dram_read_bytes': rdb, 'dram_write_bytes': wrb,
           'algorithmic_bytes_per_launch': algorithmic_bytes, 'metrics': m, 'source': rep,
           'note': 'ncu --set full --clock-control none --import-source on, one launch'}
    json.dump(res, open(out, 'w'), indent=1)
    print(json.dumps({k: res[k] for k in ('kernel', 'duration_ms_under_ncu', 'dram_bytes_per_launch')}))


if __name__ == '__main__':
    if sys.argv[1] == 'launches':
        launches(sys.argv[2], sys.argv[3], sys.argv[4] if len(sys.argv) > 4 else 'ncu launch list')
    else:
        kernel(sys.argv[2], sys.argv[3], sys.argv[4], int(sys.argv[5]) if len(sys.argv) > 5 else None)
