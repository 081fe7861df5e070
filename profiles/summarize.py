"""Turns the raw ncu outputs brought back in gpurun_out/ into the committed summaries under profiles/.

    python profiles/summarize.py <round-tag> [launches.csv] [report.ncu-rep] [kernel]

Writes profiles/<tag>_launches.md (every kernel of one bench step with its device time and share),
profiles/<tag>_<kernel>_metrics.csv (the ncu --set full metrics that matter, per captured launch) and
profiles/<kernel>_traffic.json (DRAM bytes per launch, read by bench.py for roofline.traffic).  `kernel` is the
kernel the report captured: k_post (default, the dominant one) or k_scan.
"""
import collections
import csv
import json
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)

METRICS = [
    'gpu__time_duration.sum', 'launch__grid_size', 'launch__block_size', 'launch__registers_per_thread',
    'launch__shared_mem_per_block_dynamic', 'launch__occupancy_limit_registers', 'launch__occupancy_limit_shared_mem',
    'sm__warps_active.avg.pct_of_peak_sustained_active', 'smsp__inst_executed.sum',
    'sm__inst_issued.avg.pct_of_peak_sustained_active', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
    'smsp__thread_inst_executed_per_inst_executed.ratio',
    'sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active',
    'sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_tc.avg.pct_of_peak_sustained_active',
    'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active',
    'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed',
    'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum', 'lts__t_bytes.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
    'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
    'smsp__average_warps_issue_stalled_wait_per_issue_active.ratio',
    'smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio',
    'smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio',
    'smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio',
    'smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio',
    'smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio',
    'smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio',
    'smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio',
]


SOURCES = {'k_post': ('ds_topn.cu', 'ds_common.cuh'), 'k_scan': ('ds_topn.cu', 'ds_common.cuh'), 'k_select': ('ds_topn.cu', 'ds_common.cuh'),
           'k_indel_pairs': ('ds_pairs.cu', 'ds_common.cuh'), 'k_feature_words': ('ds_pairs.cu', 'ds_common.cuh'),
           'k_query_pairs': ('ds_pairs.cu', 'ds_common.cuh'), 'k_query_features': ('ds_pairs.cu', 'ds_common.cuh')}
WORKLOAD = os.environ.get('DS_PROFILE_WORKLOAD', 'bench.py --steps 1 --warmup 1 --no-cpu (C3: 100k x 500k synthetic, top-10)')


def to_ms(value, unit):
    value = float(value.replace(',', ''))
    return {'ns': 1e-6, 'nsecond': 1e-6, 'us': 1e-3, 'usecond': 1e-3, 'ms': 1.0, 'msecond': 1.0, 's': 1e3}.get(unit, 1e-6) * value


def to_bytes(value, unit):
    value = float(value.replace(',', ''))
    return value * {'byte': 1, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9}.get(unit, 1)


def launches(tag, path):
    lines = [l for l in open(path) if not l.startswith('==')]
    agg = collections.OrderedDict()
    total = 0.0
    for row in csv.DictReader(lines):
        name = row['Kernel Name'].split('(')[0].replace('void ', '')
        ms = to_ms(row['Metric Value'], row['Metric Unit'])
        entry = agg.setdefault(name, [0, 0.0, 0.0])
        entry[0] += 1
        entry[1] += ms
        entry[2] = max(entry[2], ms)
        total += ms
    out = [f'# {tag}: kernel launch list of `python bench.py --steps 1 --warmup 1 --no-cpu`', '',
           'ncu --metrics gpu__time_duration.sum --clock-control none (per-launch times are cold-cache and serialised:',
           'compare SHARES).  The list covers the index build, one warm-up step, one timed device step and the e2e steps.', '',
           '| kernel | launches | total ms | share | longest ms |', '|---|---:|---:|---:|---:|']
    for name, (count, ms, longest) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        out.append(f'| `{name[:70]}` | {count} | {ms:.3f} | {ms / total:.3f} | {longest:.3f} |')
    out.append(f'| **total** | {sum(v[0] for v in agg.values())} | {total:.3f} | 1.000 | |')
    open(os.path.join(HERE, f'{tag}_launches.md'), 'w').write('\n'.join(out) + '\n')


def metrics(tag, report, kernel):
    raw = subprocess.run(['ncu', '-i', report, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    header, units = rows[0], rows[1]
    picked = [m for m in METRICS if m in header]
    with open(os.path.join(HERE, f'{tag}_{kernel}_metrics.csv'), 'w', newline='') as f:
        w = csv.writer(f)
        w.writerow(['metric', 'unit'] + [f'launch_{i}' for i in range(len(rows) - 2)])
        for m in picked:
            i = header.index(m)
            w.writerow([m, units[i]] + [r[i] for r in rows[2:]])
    dram = []
    for r in rows[2:]:
        total = 0.0
        for m in ('dram__bytes_read.sum', 'dram__bytes_write.sum'):
            i = header.index(m)
            total += to_bytes(r[i], units[i])
        dram.append(total)
    json.dump({'kernel': kernel, 'dram_bytes_per_launch': sum(dram) / len(dram), 'captured_launches': dram,
               'source': f'profiles/{tag}_{kernel}_metrics.csv (ncu --set full --clock-control none)'},
              open(os.path.join(HERE, f'{kernel}_traffic.json'), 'w'), indent=1)

    def mean_of(metric, scale=1.0):
        if metric not in header:
            return None
        i = header.index(metric)
        values = [float(r[i].replace(',', '')) for r in rows[2:] if r[i] not in ('', 'n/a')]
        return sum(values) / len(values) * scale if values else None

    def bytes_of(metric):
        if metric not in header:
            return None
        i = header.index(metric)
        values = [to_bytes(r[i], units[i]) for r in rows[2:] if r[i] not in ('', 'n/a')]
        return sum(values) / len(values) if values else None
    i_time = header.index('gpu__time_duration.sum')
    # what bench.py reads for `roofline` (the measured limiter of the kernel), stamped with the kernel sources it is from
    import hashlib
    digest = hashlib.sha256()
    for name in SOURCES.get(kernel, ('ds_topn.cu', 'ds_common.cuh')):
        with open(os.path.join(ROOT, 'doppelspeller_b200', 'csrc', name), 'rb') as f:
            digest.update(f.read())
    json.dump({'kernel': kernel,
               'smem_wavefronts_pct': mean_of('l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed'),
               'issue_active_pct': mean_of('smsp__issue_active.avg.pct_of_peak_sustained_active'),
               'dram_throughput_pct': mean_of('gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed'),
               'warps_active_pct': mean_of('sm__warps_active.avg.pct_of_peak_sustained_active'),
               'lsu_pipe_pct': mean_of('sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active'),
               'alu_pipe_pct': mean_of('sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active'),
               'l2_bytes_per_launch': bytes_of('lts__t_bytes.sum'), 'dram_bytes_per_launch': sum(dram) / len(dram),
               'captured_launch_ms': sum(to_ms(r[i_time], units[i_time]) for r in rows[2:]) / len(rows[2:]),
               'registers': mean_of('launch__registers_per_thread'), 'workload': WORKLOAD,
               'source': f'profiles/{tag}_{kernel}_metrics.csv (ncu --set full --clock-control none --import-source on)',
               'kernel_sources': digest.hexdigest()[:16]},
              open(os.path.join(HERE, f'r2_{kernel}_ncu.json'), 'w'), indent=1)


if __name__ == '__main__':
    tag = sys.argv[1]
    launches(tag, sys.argv[2] if len(sys.argv) > 2 else os.path.join(ROOT, 'gpurun_out', 'launches_r1.csv'))
    metrics(tag, sys.argv[3] if len(sys.argv) > 3 else os.path.join(ROOT, 'gpurun_out', 'prof_scan_r1.ncu-rep'),
            sys.argv[4] if len(sys.argv) > 4 else 'k_post')
