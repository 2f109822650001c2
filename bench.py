#!/usr/bin/env python
"""bench.py -- headline benchmark of the Telogator2 hot path on B200.

One "step" = one pass of the hot path over one batch of synthetic input derived from the bundled
fixtures (tests/golden):
  * DP   : the iter-0 all-vs-all TVR distance matrix (tvr matrix, gaps -4, adjust_lens + min_viable,
           R = 2, seed 12345) over N colour vectors of length 1000 (SURVEY 8d "D1"),
  * scan : base counts + orientation + density x2 + non-overlapping hits on the last 6000 bases
           (SURVEY 8d "S x r") over replicated reads with seeded edits.
`value` is DP GCUPS (cells as defined in SURVEY 8d / BASELINE.md 3) with inputs resident in HBM;
`e2e` is the same metric through the public API (host strings in, float32 matrix out).  The scan
numbers ride along in the `scan` object.  `--impl reference` times the CPU oracle port
(oracle/tg_oracle.c, OpenMP over all host cores) on a bounded sample of the same workload.
"""
import argparse
import gzip
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, 'tests', 'golden')
METRIC = 'tvr_all_vs_all_gcups'
PARITY_NOTE = ('bit-exact vs the CPU oracle; the oracle is pinned to the reference for the scan, the RNG, the distance epilogue, the '
               'batch boundaries and the colour-vector preparation; Bio.Align.PairwiseAligner.score itself is UNPINNED (Biopython is '
               'installable neither in the build image nor on the GPU box: profiles/r02_pin_biopython_*.log); two independent '
               'restatements of its Needleman-Wunsch score path agree (tests/test_oracle_golden.py)')
LETTER_MIX = 'CCCCCCCCCCCCVVVHTFDA'


def workload_config(n_seq, world, scan_mbases_total, scaling='weak'):
    n_pairs = n_seq * (n_seq - 1) // 2
    return {'workload': f'D1 iter-0 TVR distance matrix (SURVEY 8d): N={n_seq} colour vectors (bundled HiFi+ONT reads, truncate 1000, replicated '
                        f'with 2% seeded letter edits), tvr matrix, gaps -4, adjust_lens+min_viable, R=2, rng_seed 12345; pair space sharded '
                        f'over {world} GPU(s) (' + ('weak scaling: N = n_seq*sqrt(GPUs)' if scaling == 'weak' else 'strong scaling: fixed N') + ')',
            'n_seq': n_seq, 'pairs': n_pairs,
            'l2_policy': 'DP inputs are register/shared-memory resident (integer-ALU bound, DRAM < 2 MB per launch); every step regenerates '
                         'all shuffles and rewrites all per-pair outputs; the scan inputs (>= 150 MB packed + 800 MB outputs) exceed the 126 MB L2',
            'scan_workload': f'{scan_mbases_total:.0f} Mbases of bundled HiFi+ONT reads replicated with 0.5% seeded edits, K=57 (+57 reverse '
                             f'complements), window 100, hits on the last 6000 bases of every read'}


def load_fixtures():
    with gzip.open(os.path.join(GOLDEN, 'colorvecs.json.gz'), 'rt') as f:
        cv = json.load(f)['colorvecs']
    with gzip.open(os.path.join(GOLDEN, 'reads.json.gz'), 'rt') as f:
        reads = json.load(f)
    tables = json.load(open(os.path.join(GOLDEN, 'kmer_tables.json')))
    return cv, reads, tables


def make_prep_workload(n_reads):
    """Colour-vector preparation: the k-mer hits of the bundled reads (tests/golden/tvr_prep_golden.json.gz,
    produced by the reference), replicated to n_reads reads in the reference's kmer_dat layout."""
    with gzip.open(os.path.join(GOLDEN, 'tvr_prep_golden.json.gz'), 'rt') as f:
        S = json.load(f)['sets']['human_hifi']
    K = len(S['meta_letters'])
    base = []
    for kd in S['kmer_dat']:
        if kd['tlen'] < 1000:
            continue
        hits = [[] for _ in range(K)]
        for k, a, b in kd['hits']:
            hits[k].append([a, b])
        base.append([hits, kd['tlen'], 0, 'FWD', kd['name'], 60, ''])
    kmer_dat = [base[i % len(base)] for i in range(n_reads)]
    meta = [[f'k{i}' for i in range(K)], [None] * K, S['meta_letters'], S['meta_flags']]
    flat = [{'hits': [(k, sp[0], sp[1]) for k, spans in enumerate(kd[0]) for sp in spans], 'tlen': kd[1]} for kd in base]
    return kmer_dat, meta, flat


def make_dp_workload(cv, n_seq, seed=12345):
    """D1: bundled colour vectors (truncate 1000) replicated with seeded 2 % letter edits."""
    rng = np.random.default_rng(seed)
    base = [c[:1000] for c in (cv['human_hifi'] + cv['human_ont']) if len(c) >= 1000]
    mix = np.frombuffer(LETTER_MIX.encode(), dtype=np.uint8)
    out = []
    for i in range(n_seq):
        s = np.frombuffer(base[i % len(base)].encode(), dtype=np.uint8).copy()
        pos = rng.choice(len(s), 20, replace=False)
        s[pos] = mix[rng.integers(len(mix), size=20)]
        out.append(s.tobytes().decode())
    return out


def make_scan_workload(reads, n_bases, seed=12345):
    """S x r: bundled HiFi + ONT reads replicated with seeded 0.5 % substitutions."""
    rng = np.random.default_rng(seed)
    base = [r for _, r in reads['human_hifi']] + [r for _, r in reads['human_ont']]
    acgt = np.frombuffer(b'ACGT', dtype=np.uint8)
    out, tot, i = [], 0, 0
    while tot < n_bases:
        s = np.frombuffer(base[i % len(base)].encode(), dtype=np.uint8).copy()
        k = max(len(s) // 200, 1)
        s[rng.integers(len(s), size=k)] = acgt[rng.integers(4, size=k)]
        out.append(s.tobytes().decode())
        tot += len(s)
        i += 1
    return out


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region."""
    Q = 'clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,' \
        'clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap'

    def __init__(self, gpu_index):
        self.idx, self.rows, self.proc = gpu_index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(['nvidia-smi', '-i', str(self.idx), '--query-gpu=' + self.Q,
                                          '--format=csv,noheader,nounits', '-lms', '200'], stdout=subprocess.PIPE, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:  # noqa: BLE001
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(',')])

    def stop(self):
        if self.proc is None:
            return {'sm_mhz': None, 'sm_max_mhz': None, 'reasons': ['nvidia-smi unavailable']}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:  # noqa: BLE001
            self.proc.kill()
        sm = [float(r[0]) for r in self.rows if len(r) >= 6 and r[0].replace('.', '').isdigit()]
        mx = [float(r[1]) for r in self.rows if len(r) >= 6 and r[1].replace('.', '').isdigit()]
        names = ['hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap']
        reasons = sorted({names[k] for r in self.rows if len(r) >= 6 for k in range(4) if r[2 + k].lower().startswith('active')})
        return {'sm_mhz': float(np.median(sm)) if sm else None, 'sm_max_mhz': max(mx) if mx else None, 'reasons': reasons,
                'samples': len(sm)}


def ncu_traffic(kernel, units=None):
    """dram read+write bytes per launch from the committed `ncu --set full` capture (profiles/r02_traffic.json).
    Captures are taken on a smaller launch; for the scan the per-base figure is scaled to this launch."""
    p = os.path.join(ROOT, 'profiles', 'r02_traffic.json')
    if not os.path.exists(p):
        return None, None
    e = json.load(open(p)).get(kernel)
    if e is None:
        return None, None
    if units is not None and 'bytes_per_base' in e:
        return e['bytes_per_base'] * units, f"{e['bytes_per_base']:.3f} B/base measured by ncu on {e['launch']}, scaled to this launch"
    return e['bytes_per_launch'], 'ncu capture of ' + e['launch']


def peaks():
    p = os.path.join(ROOT, 'MEASURED_PEAKS.json')
    if os.path.exists(p):
        return json.load(open(p)), 'measured (MEASURED_PEAKS.json)'
    return {'hbm_gbs': 6650.0}, 'fallback (B200_PROFILING.md)'


# ----------------------------------------------------------------------------------------------------
# reference arm: the CPU oracle port, all host threads, bounded sample
# ----------------------------------------------------------------------------------------------------
def cpu_dp_sample(seqs, n_pairs_sample, threads):
    import oracle
    P = oracle.AlignerParams(oracle.scoring_matrix('C', 'tvr'), (True, True))
    t0 = time.perf_counter()
    _D, _a, _r, cells = oracle.dist_matrix_c(seqs, P, True, True, 2, 12345, pair_range=(0, n_pairs_sample),
                                             n_threads=threads, want_scores=True)
    dt = time.perf_counter() - t0
    return cells / dt / 1e9, dt, cells


def cpu_scan_sample(reads, tables, threads, budget_bases):
    import oracle
    from concurrent.futures import ThreadPoolExecutor
    T = tables['human_hifi']
    rc = oracle.rc
    KL, CANON = T['KMER_LIST'], T['CANONICAL_STRINGS']
    KLR, CR = [rc(k) for k in KL], [rc(k) for k in CANON]
    ISSUB = T['KMER_ISSUBSTRING']
    sample, tot = [], 0
    for r in reads:
        if tot >= budget_bases:
            break
        sample.append(r)
        tot += len(r)

    def one(r):                                  # ctypes releases the GIL inside the C calls
        _f, ori, _c0, _c1, _bc = oracle.scan_read(r, KL, KLR, CANON, CR, 100)
        oracle.nonoverlapping_hits(ori[-6000:], KLR, ISSUB)
    t0 = time.perf_counter()
    with ThreadPoolExecutor(max_workers=threads) as ex:
        list(ex.map(one, sample))
    dt = time.perf_counter() - t0
    return tot / dt / 1e6, dt, tot


def scan_gtrc_params(T, rc):
    KL, CANON = T['KMER_LIST'], T['CANONICAL_STRINGS']
    return ['hifi', 60, KL, [rc(k) for k in KL], T['KMER_ISSUBSTRING'], CANON, [rc(k) for k in CANON], 100, 0.5, 100, 400, 100, 1000,
            False, '', 100, 1000]


def cpu_scan_e2e_sample(reads, tables, threads, budget_bases):
    """CPU arm of `scan_e2e`: gtrc_parallel_job (tg_tel.py:238-328) per read through the oracle's C scan functions and its
    numpy boundary calling (oracle/boundary.py), host threads over reads."""
    import oracle
    from oracle import boundary
    from concurrent.futures import ThreadPoolExecutor
    P = scan_gtrc_params(tables['human_hifi'], oracle.rc)
    [_rt, _mq, KL, KLR, ISSUB, CANON, CR, W, THR, MINS, F_TEL, F_NONTEL, F_SUB, _p, _d, _mi, _ma] = P
    sample, tot = [], 0
    for r in reads:
        if tot >= budget_bases:
            break
        sample.append(r)
        tot += len(r)

    def one(r):
        _f, ori, c0, c1, _bc = oracle.scan_read(r, KL, KLR, CANON, CR, W)
        (tl, nontel, _inter) = boundary.get_terminating_tl(c0.astype(np.float64) / W, c1.astype(np.float64) / W, 'q', W, THR,
                                                           min(F_TEL, MINS))
        tl, nontel = min(tl, len(ori)), min(nontel, len(ori))
        if not (tl < F_TEL or nontel > F_NONTEL or len(ori) < tl + F_SUB):
            return oracle.nonoverlapping_hits(ori[max(len(ori) - tl - F_SUB, 0):], KLR, ISSUB)
        return None
    t0 = time.perf_counter()
    with ThreadPoolExecutor(max_workers=threads) as ex:
        kept = sum(1 for h in ex.map(one, sample) if h is not None)
    dt = time.perf_counter() - t0
    return tot / dt / 1e6, dt, tot, kept


_JSON_FD = None


def claim_stdout():
    """The contract is ONE JSON line on stdout.  Libraries print banners there (e.g. 'NCCL version ...' on the first collective), so
    everything written to file descriptor 1 from now on goes to stderr and emit() writes the line to the real stdout."""
    global _JSON_FD
    if _JSON_FD is None:
        sys.stdout.flush()
        _JSON_FD = os.dup(1)
        os.dup2(2, 1)


def emit(line):
    data = (json.dumps(line) + '\n').encode()
    sys.stdout.flush()
    os.write(_JSON_FD if _JSON_FD is not None else 1, data)


def run_reference(args, rank, world):
    if rank != 0:
        return
    cv, reads, tables = load_fixtures()
    threads = os.cpu_count() or 1
    n_seq = int(round(args.n_seq * np.sqrt(world))) if args.scaling == 'weak' else int(args.n_seq)
    seqs = make_dp_workload(cv, n_seq)
    sample_pairs = max(64, int(args.cpu_pairs))
    vals, times = [], []
    for it in range(args.warmup + args.steps):
        g, dt, cells = cpu_dp_sample(seqs, sample_pairs, threads)
        if it >= args.warmup:
            vals.append(g)
            times.append(dt)
    scan_reads = make_scan_workload(reads, int(args.cpu_scan_mbases * 1e6))
    mb, sdt, sb = cpu_scan_sample(scan_reads, tables, threads, int(args.cpu_scan_mbases * 1e6))
    emb, edt, eb, ekept = cpu_scan_e2e_sample(scan_reads, tables, threads, int(args.cpu_scan_mbases * 1e6))
    v = float(np.mean(vals))
    line = {'impl': 'reference', 'metric': METRIC, 'value': v, 'unit': 'GCUPS', 'n_gpus': args.gpus, 'steps': args.steps,
            'warmup': args.warmup, 'ms_per_step': float(np.mean(times) * 1e3), 'higher_is_better': True, 'scaling': args.scaling,
            'vs_baseline': None, 'dtype': 'f64', 'data': 'synthetic',
            'config': workload_config(n_seq, world, args.scan_mbases * world, args.scaling),
            'cpu_baseline': {'value': v, 'unit': 'GCUPS', 'cores': threads, 'kind': 'port',
                             'sample': f'oracle/tg_oracle.c (Biopython-style double-precision NW + MT19937 shuffles), OpenMP, first '
                                       f'{sample_pairs} pairs of the same matrix per step; Biopython itself is not installable here'},
            'e2e': {'value': v, 'unit': 'GCUPS', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
            'scan': {'value': mb, 'unit': 'Mbases/s', 'sample': f'{sb} bases, oracle C scan (base count x2, density x2, hits on last 6000)',
                     'cores': threads},
            'scan_e2e': {'value': emb, 'unit': 'Mbases/s', 'cores': threads, 'reads_kept': ekept,
                         'sample': f'{eb} bases ({edt:.1f} s): gtrc_parallel_job per read through the oracle C scan, host threads over reads, '
                                   f'numpy boundary calling; the reference Python itself runs 0.56-1.0 Mbases/s/core (SURVEY 3)'},
            'parity': PARITY_NOTE,
            'gpu_launches': 0}
    emit(line)


# ----------------------------------------------------------------------------------------------------
# our arm
# ----------------------------------------------------------------------------------------------------
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=3)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--impl', default='b200', choices=['b200', 'reference'])
    ap.add_argument('--n-seq', type=int, default=2048, help='sequences per GPU in the DP workload (weak scaling)')
    ap.add_argument('--scan-mbases', type=float, default=400.0, help='scan workload per GPU, Mbases')
    ap.add_argument('--cpu-pairs', type=int, default=30000, help='pairs per step of the CPU sample (~10 s of 16 cores)')
    ap.add_argument('--cpu-scan-mbases', type=float, default=200.0, help='bases of the CPU scan sample, Mbases')
    ap.add_argument('--no-cpu-baseline', action='store_true')
    ap.add_argument('--scaling', default='weak', choices=['weak', 'strong'],
                    help='weak: N = n_seq * sqrt(GPUs) (pairs per GPU fixed); strong: the same N = n_seq matrix on every GPU count')
    args = ap.parse_args()
    claim_stdout()
    rank = int(os.environ.get('RANK', '0'))
    world = int(os.environ.get('WORLD_SIZE', '1'))
    local_rank = int(os.environ.get('LOCAL_RANK', '0'))
    if args.impl == 'reference':
        run_reference(args, rank, world)
        return

    import torch
    import torch.distributed as dist
    assert torch.cuda.is_available(), 'bench.py needs a CUDA device (no CPU fallback)'
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group('nccl', device_id=torch.device('cuda', local_rank))
    from telogator2_b200 import tg_align, tg_kmer
    from telogator2_b200.dist_engine import DistEngine, Scoring
    from telogator2_b200.multi_gpu import local_only, shard_pair_indices, reduce_matrix, sharded_dist_matrix

    cv, reads, tables = load_fixtures()
    # weak scaling: the pair space grows with the number of GPUs (n_seq * sqrt(world)), each rank owns 1/world of it
    n_seq = int(round(args.n_seq * np.sqrt(world))) if args.scaling == 'weak' else int(args.n_seq)
    seqs = make_dp_workload(cv, n_seq)
    n_pairs = n_seq * (n_seq - 1) // 2
    aligner = tg_align.get_aligner_object(tg_align.get_scoring_matrix('C', 'tvr'), (True, True), 'tvr')
    R, SEED = 2, 12345
    my_pairs = shard_pair_indices(n_pairs, rank, world)

    T = tables['human_hifi']
    KL, CANON, ISSUB = T['KMER_LIST'], T['CANONICAL_STRINGS'], T['KMER_ISSUBSTRING']
    KLR, CR = [tg_kmer.RC(k) for k in KL], [tg_kmer.RC(k) for k in CANON]
    scan_reads = make_scan_workload(reads, int(args.scan_mbases * 1e6), seed=12345 + rank)
    scan_bases = sum(len(r) for r in scan_reads)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---------------- DP: device-resident timing ----------------
    # the encoded sequences are resident in HBM before the timed region; a step = every kernel of every batch of this
    # rank's pair ranges + the float epilogue into the device matrix (+ the matrix all_reduce over NVLink when world > 1)
    eng = DistEngine(Scoring.from_aligner(aligner))
    dp_state = eng.prepare(seqs)
    assert dp_state is not None

    def dp_step():
        eng.stats['cells'] = 0
        eng.stats['launches'] = 0
        D = eng.matrix_device(dp_state, True, True, R, SEED, shard=(rank, world) if world > 1 else None)
        if world > 1:
            reduce_matrix(D, dist)
        return D

    # ---------------- scan: device-resident timing ----------------
    pr = tg_kmer.PackedReads(scan_reads)
    lens = pr.lens
    sl_idx = np.arange(pr.n)
    sl_lo = np.maximum(lens - 6000, 0)

    orig_planes = (pr.d_pack2, pr.d_nmask)
    sl_hi = lens.copy()

    scan_ev = [torch.cuda.Event(enable_timing=True) for _ in range(6)]
    scan_parts = {'basecount': [], 'orient': [], 'density': [], 'hits': []}

    def scan_step(record=False):
        pr.d_pack2, pr.d_nmask = orig_planes
        scan_ev[0].record()
        bc = pr.basecount(CANON, CR)                               # tg_tel.py:262-263
        scan_ev[1].record()
        flip = (bc[:, 0] > bc[:, 1]).cpu().numpy()
        scan_ev[2].record()
        pr.orient(flip)                                            # tg_tel.py:265-266
        scan_ev[3].record()
        ca, cb = pr.density_counts(KL, KLR, 100)                   # tg_tel.py:160-161
        scan_ev[4].record()
        spans = pr.hits_device(sl_idx, sl_lo, sl_hi, KLR, ISSUB)   # tg_tel.py:310 on the last 6000 bases
        scan_ev[5].record()
        if record:
            torch.cuda.synchronize()
            for name, a, b in (('basecount', 0, 1), ('orient', 2, 3), ('density', 3, 4), ('hits', 4, 5)):
                scan_parts[name].append(scan_ev[a].elapsed_time(scan_ev[b]))
        return flip, ca, cb, spans

    import ctypes as C
    from telogator2_b200 import _lib
    L = _lib.load()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
    sampler = ClockSampler(local_rank)
    dp_ms, scan_ms, dens_ms = [], [], []
    cells_step = 0
    for it in range(args.warmup + args.steps):
        timed = it >= args.warmup
        if it == args.warmup:
            barrier()
            sampler.start()
            t_wall0 = time.perf_counter()
            _lib.check(L.tg_wave_timing(1))            # CUDA events around every nw_wave_kernel launch of the timed steps
        barrier()
        ev[0].record()
        dp_step()
        ev[1].record()
        torch.cuda.synchronize()
        cells_step = eng.stats['cells']
        launches_dp = eng.stats['launches']
        ev[2].record()
        scan_step(record=timed)
        ev[3].record()
        barrier()
        if timed:
            dp_ms.append(ev[0].elapsed_time(ev[1]))
            scan_ms.append(ev[2].elapsed_time(ev[3]))
    t_wall = time.perf_counter() - t_wall0
    clocks = sampler.stop()
    wave_ms, wave_n = C.c_double(0), C.c_int(0)
    _lib.check(L.tg_wave_timing_read(C.byref(wave_ms), C.byref(wave_n)))
    _lib.check(L.tg_wave_timing(0))

    # boundary calling on the device-resident window counts of the last step (tg_kmer.py:105-170, tg_tel.py:164-233)
    boundary = None
    try:
        from telogator2_b200 import tg_boundary
        _f, ca_, cb_, _s = scan_step()
        del _s
        bt = []
        for _ in range(3):
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            term = tg_boundary.terminating_tl_from_counts(ca_, cb_, pr.d_base_off, pr.lens, 100, 0.5, 100, 'q')
            torch.cuda.synchronize()
            bt.append(time.perf_counter() - t0)
        boundary = {'ms': min(bt) * 1e3, 'mbases_per_s': float(pr.lens.sum()) / min(bt) / 1e6, 'reads_with_telomere': int((term[:, 0] > 0).sum()),
                    'what': 'power signal + db4 wavelet smoothing + region walk + scoring for every read, host wall clock incl. the '
                            '16 B/read result copy; ~100 B of float64 traffic per base'}
        del ca_, cb_
    except Exception as exc:                          # an auxiliary line must not cost the headline metric
        boundary = {'error': f'{type(exc).__name__}: {exc}'}

    def allmax(x):
        t = torch.tensor([x], dtype=torch.float64, device='cuda')
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def allsum(x):
        t = torch.tensor([x], dtype=torch.float64, device='cuda')
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.SUM)
        return float(t.item())

    dp_t = allmax(float(np.mean(dp_ms)) * 1e-3)
    scan_t = allmax(float(np.mean(scan_ms)) * 1e-3)
    cells_all = allsum(float(cells_step))
    bases_all = allsum(float(scan_bases))
    gcups = cells_all / dp_t / 1e9
    scan_mbs = bases_all / scan_t / 1e6

    # ---------------- kernel-level roofline numbers (rank 0, CUDA events around single launches) -----
    roof, roof_scan = None, None
    if rank == 0:
        dpx, mixed = C.c_double(0), C.c_double(0)
        _lib.check(L.tg_int_alu_peak(1000, C.byref(dpx), C.byref(mixed)))
        n_sm = C.c_int(0)
        _lib.check(L.tg_device_sm_count(C.byref(n_sm)))
        nominal = n_sm.value * 64 * float(clocks.get('sm_max_mhz') or 1965.0) * 1e6     # ALU pipe: 64 lanes per SM per clock
        # dominant kernel: nw_wave_kernel, timed inside the timed steps (CUDA events on the launching stream around every
        # launch: phase A + phase B of every batch).  Algorithmic work = the cells of the steps.
        k_cells = float(cells_step) * args.steps
        k_gcups = k_cells / (wave_ms.value * 1e-3) / 1e9
        instr_per_cell = 1.5                       # minimal s16x2 cost of the linear-gap cell
        peak_gcups = dpx.value / instr_per_cell / 1e9
        roof = {'bound': 'int_alu', 'kernel': 'nw_wave_kernel', 'achieved': k_gcups, 'peak': peak_gcups, 'unit': 'GCUPS',
                'frac': k_gcups / peak_gcups, 'traffic': ncu_traffic('nw_wave_kernel')[0], 'traffic_how': ncu_traffic('nw_wave_kernel')[1],
                'peak_source': f'measured in this run: {dpx.value / 1e12:.2f} T VIMNMX3.U16x2 lane-ops/s (tg_int_alu_peak: '
                               f'{100 * dpx.value / nominal:.1f} % of SMs x 64 lanes x max clock) / 1.5 instr per cell '
                               f'(SURVEY 8d); add+max3 pair rate {mixed.value / 1e12:.2f} T pairs/s',
                'launch_ms': wave_ms.value / max(wave_n.value, 1), 'launches': wave_n.value,
                'cells_per_launch': k_cells / max(wave_n.value, 1),
                'share_of_step': wave_ms.value / (float(np.sum(dp_ms)) if dp_ms else 1.0),
                'how': 'CUDA events around every nw_wave_kernel launch (phase A and phase B of every batch) of the timed steps on rank 0'}
        # scan: density kernel alone, algorithmic bytes 2.375 B/base
        d0, d1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ds = []
        for _ in range(3):
            d0.record()
            pr.density_counts(KL, KLR, 100)
            d1.record()
            torch.cuda.synchronize()
            ds.append(d0.elapsed_time(d1) * 1e-3)
        pk, src = peaks()
        ach = scan_bases * 2.375 / min(ds) / 1e9
        roof_scan = {'bound': 'hbm', 'kernel': 'scan_cov_kernel<density>', 'achieved': ach, 'peak': pk['hbm_gbs'], 'unit': 'GB/s',
                     'frac': ach / pk['hbm_gbs'], 'traffic': ncu_traffic('scan_cov_kernel<density>', scan_bases)[0],
                     'traffic_how': ncu_traffic('scan_cov_kernel<density>', scan_bases)[1], 'peak_source': src, 'launch_ms': min(ds) * 1e3,
                     'bytes_per_base': 2.375, 'mbases_per_s': scan_bases / min(ds) / 1e6}

    # ---------------- D2: many small matrices, iteration-2 shape (N=48, L<=3000, R=3) ----------------
    # (single-process runs only: the public API it calls is collective under a process group, and the other ranks are
    #  not taking part here)
    d2 = None
    if world == 1:
        try:
            rs2 = np.random.default_rng(21)
            long_cv = [c for c in (cv['human_ont'] + cv['human_hifi']) if len(c) >= 3000]
            groups = []
            for gi in range(64):
                grp = []
                for k in range(48):
                    sq = list(long_cv[(gi * 48 + k) % len(long_cv)][:3000])
                    for ppos in rs2.choice(3000, 60, replace=False):
                        sq[ppos] = 'CVHTFDA'[rs2.integers(7)]
                    grp.append(''.join(sq))
                groups.append(grp)
            tg_align.get_dist_matrices_parallel(groups[:2], aligner, True, True, 3, rng_seed=SEED)
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            many = tg_align.get_dist_matrices_parallel(groups, aligner, True, True, 3, rng_seed=SEED)
            t_many = time.perf_counter() - t0
            t0 = time.perf_counter()
            ones = [tg_align.get_dist_matrix_parallel(gq, aligner, True, True, 3, rng_seed=SEED) for gq in groups[:16]]
            t_one = (time.perf_counter() - t0) * len(groups) / 16
            assert all(np.array_equal(a, b) for a, b in zip(many[:16], ones)), 'batched matrices differ from the per-call matrices'
            e3 = DistEngine(Scoring.from_aligner(aligner))
            e3.matrix_device(e3.prepare(groups, many=True), True, True, 3, SEED)
            d2 = {'matrices': len(groups), 'n_seq': 48, 'len': 3000, 'R': 3, 'cells': int(e3.stats['cells']),
                  'batched_gcups': e3.stats['cells'] / t_many / 1e9, 'batched_ms': t_many * 1e3,
                  'per_call_gcups': e3.stats['cells'] / t_one / 1e9, 'per_call_ms': t_one * 1e3,
                  'what': 'host strings -> host matrices; batched = get_dist_matrices_parallel (one pass), per_call = one '
                          'get_dist_matrix_parallel per matrix (16 timed, scaled)'}
            del many, ones, e3

        except Exception as exc:                      # an auxiliary line must not cost the headline metric
            d2 = {'error': f'{type(exc).__name__}: {exc}'}
    # ---------------- hierarchical clustering of the bench matrix ("scipy linkage itself") ----------------
    linkage_line = None
    if world == 1:
        try:
            from telogator2_b200 import tg_linkage
            from scipy.cluster.hierarchy import linkage as scipy_linkage
            from scipy.spatial.distance import squareform
            d_D = dp_step()                                   # the float32 matrix where the distance kernels leave it
            torch.cuda.synchronize()
            tg_linkage.linkage_from_matrix(d_D, 'ward', divide_by=10.0)
            lt = []
            for _ in range(3):
                torch.cuda.synchronize()
                t0 = time.perf_counter()
                Zg = tg_linkage.linkage_from_matrix(d_D, 'ward', divide_by=10.0)
                lt.append(time.perf_counter() - t0)
            Mh = d_D.cpu().numpy()
            t0 = time.perf_counter()
            Mh /= 10.0                                        # tg_tvr.py:154
            Zs = scipy_linkage(squareform(Mh, checks=False), method='ward')       # tg_tvr.py:166-167
            t_sc = time.perf_counter() - t0
            assert np.array_equal(Zg, Zs), 'device linkage differs from scipy'
            linkage_line = {'n': int(n_seq), 'method': 'ward', 'gpu_ms': min(lt) * 1e3, 'scipy_ms': t_sc * 1e3, 'identical_to_scipy': True,
                            'what': 'linkage(squareform(D / 10), "ward") from the device matrix (merge list to the host) vs scipy on the host matrix'}
            del d_D, Mh
        except Exception as exc:                      # an auxiliary line must not cost the headline metric
            linkage_line = {'error': f'{type(exc).__name__}: {exc}'}
    # ---------------- colour-vector preparation, host lists in -> host lists out ----------------
    prep = None
    if world == 1:
        try:
            from telogator2_b200 import tg_tvr
            kd_prep, meta_prep, flat_prep = make_prep_workload(2048)
            tg_tvr.tvr_prep(kd_prep[:64], meta_prep, 1000)
            pt = []
            for _ in range(3):
                torch.cuda.synchronize()
                t0 = time.perf_counter()
                got_prep = tg_tvr.tvr_prep(kd_prep, meta_prep, 1000)
                pt.append(time.perf_counter() - t0)
            letters = sum(kd[1] for kd in kd_prep)
            prep = {'value': len(kd_prep) / min(pt), 'unit': 'reads/s', 'ms_per_step': min(pt) * 1e3, 'reads': len(kd_prep),
                    'letters': int(letters), 'what': 'tvr_prep: spans -> colour vectors, 3 density boundaries, denoise, string assembly '
                                                     '(kmer_dat lists in, Python strings out)'}
            if not args.no_cpu_baseline:
                import oracle
                t0 = time.perf_counter()
                want = oracle.tvr_prep_py(flat_prep, meta_prep, 1000)
                dt = time.perf_counter() - t0
                assert got_prep['tvrtels_for_clustering'][:len(flat_prep)] == want['tvrtels_for_clustering'], 'tvr_prep differs from the oracle'
                prep['cpu_port_reads_per_s'] = len(flat_prep) / dt
                prep['cpu_port'] = f'oracle numpy/Python restatement, 1 core, {len(flat_prep)} reads ({dt:.2f} s)'

        except Exception as exc:                      # an auxiliary line must not cost the headline metric
            prep = {'error': f'{type(exc).__name__}: {exc}'}
    # ---------------- stage [0a] prefilter: str.count of KMER_INITIAL / RC over every read ----------------
    prefilter = None
    if world == 1:
        try:
            from telogator2_b200 import tg_prefilter
            g = torch.Generator(device='cuda')
            g.manual_seed(99)
            n_wgs, wgs_len = 131072, 16384                                 # 2.1 Gbases of genome-like reads (uniform ACGT) ...
            tail = tg_prefilter._lib.load().tg_prefilter_tail_bytes()
            d_wgs = torch.empty(n_wgs * wgs_len + tail, dtype=torch.uint8, device='cuda')
            acgt = torch.tensor([65, 67, 71, 84], dtype=torch.uint8, device='cuda')
            for c0 in range(0, d_wgs.numel(), 1 << 26):
                c1 = min(c0 + (1 << 26), d_wgs.numel())
                d_wgs[c0:c1] = acgt[torch.randint(0, 4, (c1 - c0,), generator=g, device='cuda')]
            # ... with telomeric scan reads spliced in (every 512th read: still ~5x the telomeric fraction of a real 30x genome)
            TEL_EVERY = 512
            tel = [r for r in scan_reads if len(r) <= wgs_len][:n_wgs // TEL_EVERY]
            for k, r in enumerate(tel):
                o = TEL_EVERY * k * wgs_len
                d_wgs[o:o + len(r)] = torch.frombuffer(bytearray(r.encode('latin-1')), dtype=torch.uint8).cuda()
            lens_wgs = np.full(n_wgs, wgs_len, dtype=np.int32)
            for k, r in enumerate(tel):
                lens_wgs[TEL_EVERY * k] = len(r)
            ar = tg_prefilter.AsciiReads.from_device(d_wgs, np.arange(n_wgs + 1, dtype=np.int64) * wgs_len, lens_wgs)
            KI, KIR = CANON[0] + CANON[0], CR[0] + CR[0]                   # telogator2.py:393-394
            ca, cb = ar.count_pair_device(KI, KIR)
            torch.cuda.synchronize()
            p0, p1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            ps = []
            for _ in range(5):
                p0.record()
                ca, cb = ar.count_pair_device(KI, KIR)
                p1.record()
                torch.cuda.synchronize()
                ps.append(p0.elapsed_time(p1) * 1e-3)
            pf_bases = float(lens_wgs.astype(np.int64).sum())
            pkp, srcp = peaks()
            # spot check against str.count (the reference's own code) on the telomeric reads and some random ones
            h_ca, h_cb = ca.cpu().numpy(), cb.cpu().numpy()
            for k in list(range(0, min(len(tel), 24))):
                assert h_ca[TEL_EVERY * k] == tel[k].count(KI) and h_cb[TEL_EVERY * k] == tel[k].count(KIR), 'prefilter count differs from str.count'
            for i in (1, 2, 3, 1000, n_wgs - 1):
                s_i = bytes(d_wgs[i * wgs_len:(i + 1) * wgs_len].cpu().numpy()).decode('latin-1')
                assert h_ca[i] == s_i.count(KI) and h_cb[i] == s_i.count(KIR), 'prefilter count differs from str.count'
            prefilter = {'kernel': 'prefilter_count_kernel', 'bound': 'hbm', 'achieved': pf_bases / min(ps) / 1e9, 'peak': pkp['hbm_gbs'],
                         'unit': 'GB/s', 'frac': pf_bases / min(ps) / 1e9 / pkp['hbm_gbs'], 'peak_source': srcp, 'launch_ms': min(ps) * 1e3,
                         'bytes_per_base': 1.0, 'mbases_per_s': pf_bases / min(ps) / 1e6, 'reads': n_wgs,
                         'telomeric_reads': len(tel), 'reads_kept': int(((h_ca >= 5) | (h_cb >= 5)).sum()),
                         'what': 'read.count(KMER_INITIAL), read.count(KMER_INITIAL_RC) for every read (telogator2.py:451-454), ASCII resident in HBM'}
            del d_wgs, ar

        except Exception as exc:                      # an auxiliary line must not cost the headline metric
            prefilter = {'error': f'{type(exc).__name__}: {exc}'}
    # ---------------- scan e2e: host reads in -> the 9-element result list out, through the batch boundary ----------------
    scan_e2e = None
    if world == 1:
        try:
            from telogator2_b200 import tg_tel
            gp = scan_gtrc_params(T, tg_kmer.RC)
            ard = [(f'read{i}', r, None) for i, r in enumerate(scan_reads)]
            tg_tel.get_tel_repeat_comp_parallel(ard[:64], gp)
            st = []
            res = None
            for _ in range(2):
                res = None                            # (dropping the previous pass's ~15 M result objects is the caller's business, not the call's)
                torch.cuda.synchronize()
                t0 = time.perf_counter()
                res = tg_tel.get_tel_repeat_comp_parallel(ard, gp)
                st.append(time.perf_counter() - t0)
            n_spans = sum(len(sp) for khd in res[0] for sp in khd[0])
            n_kept = len(res[0])
            del res, ard               # (millions of small lists: leaving them alive makes every later cycle collection slow)
            import gc
            gc.collect()
            scan_e2e = {'value': scan_bases / float(np.mean(st)) / 1e6, 'unit': 'Mbases/s', 'ms_per_step': float(np.mean(st)) * 1e3,
                        'reads': len(scan_reads), 'reads_kept': n_kept, 'spans': int(n_spans),
                        'h2d_bytes_per_step': int(tg_tel.LAST_STATS.get('h2d_bytes', 0)), 'd2h_bytes_per_step': int(tg_tel.LAST_STATS.get('d2h_bytes', 0)),
                        'phases_ms': {k: float(v) * 1e3 for k, v in tg_tel.LAST_STATS.get('phases', {}).items()},
                        'boundary': tg_tel.LAST_STATS.get('boundary'), 'hits_phases_ms': {k: float(v) * 1e3 for k, v in tg_tel.LAST_STATS.get('hits_phases', {}).items()},
                        'what': 'tg_tel.get_tel_repeat_comp_parallel: all_read_dat (Python strings) in -> the 9-element result (kmer_hit_dat '
                                'lists etc.) out; boundary calling (get_terminating_tl) on the device'}
        except Exception as exc:                      # an auxiliary line must not cost the headline metric
            scan_e2e = {'error': f'{type(exc).__name__}: {exc}'}
    # ---------------- e2e through the public API: host strings -> float32 matrix on the host --------
    # (the body of tg_align.get_dist_matrix_parallel, with the engine kept to read its byte counters)
    e2e_ms = []
    h2d = d2h = 0
    D = None
    n_e2e_warm = 2                             # untimed passes first (device allocator and page-locked pool warm, like `value`'s warm-up steps)
    for it in range(n_e2e_warm + max(3, min(args.steps, 5))):
        del D                                  # a caller drops the previous matrix; its page-locked block is reused
        barrier()
        t0 = time.perf_counter()
        eng2 = DistEngine(Scoring.from_aligner(aligner))
        D = sharded_dist_matrix(eng2, seqs, True, True, R, SEED)
        barrier()
        if it >= n_e2e_warm:
            e2e_ms.append((time.perf_counter() - t0) * 1e3)
        h2d, d2h = eng2.stats['h2d_bytes'], eng2.stats['d2h_bytes']
    verified = None
    if rank == 0:
        with local_only():                   # rank 0 alone: nothing in here may take part in a collective
            # cross-check the reduced multi-GPU matrix on a sample of pairs recomputed locally through the
            # explicit-job path + host epilogue (same global pair indices -> same seeds)
            rs = np.random.default_rng(7)
            sample = np.sort(rs.choice(n_pairs, size=min(256, n_pairs), replace=False)).astype(np.int64)
            chk = DistEngine(Scoring.from_aligner(aligner)).pair_scores(seqs, True, True, R, SEED, pair_idx=sample)
            want = DistEngine.distances({k: chk[k][sample] for k in ('aln', 'rnd', 'iden')}, R).astype('<f4')
            ar = np.arange(n_seq, dtype=np.int64)
            row_start = ar * (2 * n_seq - ar - 1) // 2
            si = np.minimum(np.searchsorted(row_start, sample, side='right') - 1, n_seq - 2)
            got = D[si, sample - row_start[si] + si + 1]
            assert np.array_equal(got, want), 'bench matrix differs from the explicit-job path + host epilogue on sampled pairs'
            assert np.array_equal(D, D.T)
            verified = int(len(sample))
    e2e_t = allmax(float(np.mean(e2e_ms)) * 1e-3)          # the same statistic as `value`: mean over the passes
    e2e_gcups = cells_all / e2e_t / 1e9
    assert D.shape == (n_seq, n_seq)

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        threads = os.cpu_count() or 1
        g, dt, cells = cpu_dp_sample(seqs, args.cpu_pairs, threads)
        smb, sdt, sb = cpu_scan_sample(scan_reads, tables, threads, int(args.cpu_scan_mbases * 1e6))
        emb, edt, eb, _k = cpu_scan_e2e_sample(scan_reads, tables, threads, int(args.cpu_scan_mbases * 1e6) // 2)
        if scan_e2e is not None and 'value' in scan_e2e:
            scan_e2e['cpu_port_mbases_per_s'] = emb
            scan_e2e['cpu_port'] = f'oracle C scan per read, {threads} host threads, {eb} bases ({edt:.1f} s), numpy boundary calling'
        cpu = {'value': g, 'unit': 'GCUPS', 'cores': threads, 'kind': 'port',
               'sample': f'oracle C port (Biopython-style f64 NW + MT19937), OpenMP {threads} threads, first {args.cpu_pairs} pairs of the same '
                         f'matrix ({dt:.1f} s); scan: {smb:.2f} Mbases/s on {sb} bases ({sdt:.1f} s)',
               'scan_mbases_per_s': smb}

    if rank == 0:
        line = {'metric': METRIC, 'value': gcups, 'unit': 'GCUPS', 'n_gpus': world, 'steps': args.steps, 'warmup': args.warmup,
                'ms_per_step': dp_t * 1e3, 'higher_is_better': True, 'scaling': args.scaling, 'vs_baseline': None, 'dtype': 'u16',
                'data': 'synthetic',
                'config': workload_config(n_seq, world, bases_all / 1e6, args.scaling),
                'e2e': {'value': e2e_gcups, 'unit': 'GCUPS', 'h2d_bytes_per_step': int(h2d), 'd2h_bytes_per_step': int(d2h),
                        'ms_per_step': e2e_t * 1e3},
                'gpu_launches': int(launches_dp + 4),
                'scan': {'value': scan_mbs, 'unit': 'Mbases/s', 'ms_per_step': scan_t * 1e3,
                         'what': 'basecount x2 + orient + density x2 + hits on the last 6000 bases of every read (device resident)',
                         'parts_ms': {k: float(np.mean(v)) for k, v in scan_parts.items() if v}, 'boundary': boundary},
                'roofline': roof, 'roofline_scan': roof_scan, 'cpu_baseline': cpu, 'clocks': clocks,
                'scan_e2e': scan_e2e, 'parity': PARITY_NOTE,
                'prep': prep, 'prefilter': prefilter, 'd2': d2, 'linkage': linkage_line,
                'verified_pairs': verified,
                'wall_s_timed_region': t_wall}
        emit(line)
    if world > 1:
        dist.destroy_process_group()


if __name__ == '__main__':
    main()
