#!/usr/bin/env python
"""gpurun_out/final (tools/capture_profiles.sh a + b) -> profiles/rNN_*: bench lines, launch list + shares of a DP step,
key ncu metrics of every hot kernel (last = warm launch), DRAM traffic per launch (bench.py's roofline.traffic).

    python tools/make_profile_summaries.py [r02]
"""
import collections
import csv
import json
import os
import shutil
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = os.path.join(ROOT, 'gpurun_out', 'final')
TAG = sys.argv[1] if len(sys.argv) > 1 else 'r02'
OUT = os.path.join(ROOT, 'profiles')
WANT = ['gpu__time_duration.sum', 'launch__grid_size', 'launch__block_size', 'launch__registers_per_thread',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'smsp__inst_executed.sum', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active', 'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum',
        'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'smsp__average_warp_latency_per_inst_issued.ratio',
        'smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio', 'smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio', 'smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_wait_per_issue_active.ratio', 'smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_dispatch_stall_per_issue_active.ratio', 'smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio', 'smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio',
        'smsp__thread_inst_executed_per_inst_executed.ratio']
BYTES = {'byte': 1, 'kbyte': 1e3, 'mbyte': 1e6, 'gbyte': 1e9}


def short(name):
    name = name[5:] if name.startswith('void ') else name
    return name.replace('<unnamed>::', '').split('(')[0]


def to_us(v, unit):
    v = float(v.replace(',', ''))
    return v / 1000 if unit.startswith('ns') else (v if unit.startswith('us') else v * 1000 if unit.startswith('ms') else v * 1e6)


def main():
    for a, b in [('bench_b200.json', 'bench_b200.json'), ('bench_reference.json', 'bench_reference.json'), ('pytest_gpu.log', 'pytest_gpu.log'),
                 ('launches.csv', 'launches.csv'), ('nvidia_smi.csv', 'nvidia_smi_idle.csv')]:
        shutil.copy(os.path.join(SRC, a), os.path.join(OUT, f'{TAG}_{b}'))
    # ---- launch list: shares of one DP step (N=1024: 2 batches) ----
    rows = list(csv.reader(open(os.path.join(SRC, 'launches.csv'))))
    hi = [i for i, r in enumerate(rows) if r and r[0] == 'ID'][0]
    hdr = rows[hi]
    ki, vi, ui = hdr.index('Kernel Name'), hdr.index('Metric Value'), hdr.index('Metric Unit')
    seq = [(short(r[ki]), to_us(r[vi], r[ui])) for r in rows[hi + 1:] if len(r) > vi]
    dp = [(k, u) for k, u in seq if k.startswith(('nw_wave', 'mt_', 'dist_'))]
    idx_a = [i for i, (k, _) in enumerate(dp) if k.startswith('nw_wave_kernel<3, 3')]
    s0 = idx_a[2]                                           # the timed step (after the warm-up step's two batches)
    s1, n_epi = s0, 0
    while n_epi < 2:
        n_epi += dp[s1][0] == 'dist_epilogue_kernel'
        s1 += 1
    step = dp[s0:s1]
    tot = sum(u for _, u in step)
    agg = collections.OrderedDict()
    for k, u in step:
        agg[k] = agg.get(k, 0) + u
    lines = [f'{k:40s} {u / 1000:8.2f} ms {100 * u / tot:5.1f} %' for k, u in agg.items()]
    lines.append(f'sum of the kernels of the step {tot / 1000:.2f} ms')
    scan = collections.OrderedDict()
    for k, u in seq:
        if k.startswith(('pack_', 'orient_', 'scan_cov', 'hits_', 'bd_')):
            scan.setdefault(k, []).append(u)
    with open(os.path.join(OUT, f'{TAG}_launches_summary.txt'), 'w') as f:
        f.write('Launch list (ncu --metrics gpu__time_duration.sum --clock-control none) of `bench.py --steps 1 --warmup 1 --n-seq 1024 '
                '--scan-mbases 200 --no-cpu-baseline`:\nkernels of the timed DP step (N=1024: 523 776 pairs in 2 batches), cold-cache and '
                'serialised -- compare SHARES.\n\n' + '\n'.join(lines) + '\n\nscan and boundary-calling kernels (200 Mbases, 10.8 k reads): '
                'launches, fastest, sum per pass\n')
        for k, v in scan.items():
            per_pass = sum(v) / max(len(scan.get('bd_power_kernel', [1])), 1) if k.startswith('bd_') else min(v)
            f.write(f'{k:28s} n={len(v):3d}  min {min(v):9.1f} us   per pass {per_pass:9.1f} us\n')
    print('\n'.join(lines))
    # ---- full capture ----
    rows = list(csv.reader(open(os.path.join(SRC, 'full_raw.csv'))))
    hdr, units = rows[0], rows[1]
    ki = hdr.index('Kernel Name')
    col = hdr.index
    last, order = {}, []
    for r in rows[2:]:
        k = short(r[ki])
        if k not in last:
            order.append(k)
        last[k] = r

    def dram(r):
        return sum(float(r[col(n)].replace(',', '')) * BYTES[units[col(n)].lower()] for n in ('dram__bytes_read.sum', 'dram__bytes_write.sum'))
    bd = [(short(r[ki]), r) for r in rows[2:] if short(r[ki]).startswith('bd_')]
    bd = bd[len(bd) // 2:]                                  # second (warm) pass
    bagg = collections.OrderedDict()
    for k, r in bd:
        a = bagg.setdefault(k, {'n': 0, 'us': 0.0, 'bytes': 0.0, 'inst': 0.0})
        a['n'] += 1
        a['us'] += to_us(r[col('gpu__time_duration.sum')], units[col('gpu__time_duration.sum')])
        a['bytes'] += dram(r)
        a['inst'] += float(r[col('smsp__inst_executed.sum')].replace(',', ''))
    plain = open(os.path.join(SRC, 'prof_plain.log')).read()
    with open(os.path.join(OUT, f'{TAG}_ncu_full_summary.txt'), 'w') as f:
        f.write('ncu --set full --clock-control none on `python tools/prof_run.py 256 100` (tools/capture_profiles.sh b).\nWorkload:\n' + plain +
                '\nOne block per kernel: the LAST launch of that kernel in the capture (second, warm pass).  The bd_* kernels (one launch per '
                'wavelet level)\nare summed per kernel at the end.\n\n')
        for k in order:
            if k.startswith('bd_'):
                continue
            f.write(f'=== {k}\n')
            for w in WANT:
                if w in hdr:
                    f.write(f'  {w:88s} {last[k][col(w)]:>16s} {units[col(w)]}\n')
        if bagg:
            tot = sum(a['us'] for a in bagg.values())
            f.write('\n=== boundary-calling kernels (tg_bound.cu), second pass, summed over their launches\n')
            for k, a in bagg.items():
                f.write(f"  {k:24s} launches {a['n']:3d}   {a['us']:9.1f} us ({100 * a['us'] / tot:4.1f} %)   dram {a['bytes'] / 1e6:9.1f} MB   "
                        f"warp-instr {a['inst'] / 1e6:8.1f} M\n")
            f.write(f"  total {tot:9.1f} us   dram {sum(a['bytes'] for a in bagg.values()) / 1e6:.1f} MB\n")
    nb = 100000599
    T = {'_source': f'profiles/{TAG}_ncu_full_summary.txt: ncu --set full --clock-control none on tools/prof_run.py 256 100 '
                    '(dram__bytes_read.sum + dram__bytes_write.sum per launch, last launch of each kernel)'}

    def put(key, kernels, launch, per_base=False):
        b = sum(dram(last[k]) for k in kernels if k in last)
        T[key] = {'bytes_per_launch': b, 'launch': launch}
        if per_base:
            T[key]['bytes_per_base'] = b / nb
    put('scan_cov_kernel<density>', ['scan_cov_kernel<1>'], '100.0 Mbases, 5416 reads', True)
    put('scan_cov_kernel<count>', ['scan_cov_kernel<0>'], '100.0 Mbases', True)
    put('hits_kernel', ['hits_kernel'], 'last 6000 bases of 5416 reads (32.4 Mbases of slices)')
    put('nw_wave_kernel', ['nw_wave_kernel<3, 3, 0, 0>'], 'phase A (shared column), N=256 (32640 pairs, 3.26e10 cells): DRAM traffic is noise, the kernel is integer-ALU bound')
    put('nw_wave_kernel<phase B>', ['nw_wave_kernel<3, 2, 0, 0>', 'nw_wave_kernel<3, 4, 0, 0>'], 'phase B (full + per-job profiles), N=256: reads the shuffled strings')
    for k in ('mt_stream_batch_kernel', 'mt_sample_warp_kernel', 'prefilter_count_kernel<12, 1>', 'lk_nnchain_kernel'):
        if k in last:
            put(k, [k], 'tools/prof_run.py 256 100')
    if bagg:
        b = sum(a['bytes'] for a in bagg.values())
        T['bd_* (boundary calling)'] = {'bytes_per_launch': b, 'launch': '5416 reads / 100.0 Mbases, all levels', 'bytes_per_base': b / nb}
    json.dump(T, open(os.path.join(OUT, f'{TAG}_traffic.json'), 'w'), indent=1)
    print('wrote', TAG, 'files to profiles/')


if __name__ == '__main__':
    main()
