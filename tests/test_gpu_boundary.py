"""GPU: telomere-boundary calling (tg_kmer.get_telomere_regions / wavelet_smooth, tg_tel.get_terminating_tl) against
  * the reference's OWN functions run over all bundled reads (tests/golden/boundary_golden.json.gz, tel_batch_gtt_golden.json.gz;
    PyWavelets itself replaced by oracle.boundary.fake_pywt -- its parity is unpinned), and
  * the numpy oracle (oracle/boundary.py) on seeded signals -- bit for bit: both evaluate every sum in the same order.
"""
import gzip
import hashlib
import json
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

HERE = os.path.dirname(os.path.abspath(__file__))


def gtt_cases():
    # (table, (FILT_TERM_TEL, FILT_TERM_NONTEL, FILT_TERM_SUBTEL)): telogator2.py defaults, and no filters at all
    return [('human_hifi', (400, 100, 1000)), ('human_ont', (400, 100, 1000)), ('maize_hifi', (400, 100, 1000)), ('human_hifi', (0, 0, 0))]


def gtt_params_of(T, filters, rc):
    KL, CANON = T['KMER_LIST'], T['CANONICAL_STRINGS']
    return [T['read_type'], 60, KL, [rc(k) for k in KL], T['KMER_ISSUBSTRING'], CANON, [rc(k) for k in CANON], 100, 0.5, 100,
            filters[0], filters[1], filters[2], False, '', 100, 1000]


def compact_tel_result(res):
    """The 9-element result with span lists and read sequences replaced by SHA-1 digests."""
    def h(s):
        return hashlib.sha1(s.encode('latin-1')).hexdigest()

    def spans(hits):
        tri = np.array([(k, sp[0], sp[1]) for k, lst in enumerate(hits) for sp in lst], dtype=np.int32).reshape(-1, 3)
        return [hashlib.sha1(tri.tobytes()).hexdigest(), int(len(tri))]

    def entry(e):
        return [spans(e[0]), int(e[1]), int(e[2]), e[3], e[4], int(e[5]), h(e[6])]
    return [[entry(e) for e in res[0]], [int(x) for x in res[1]], [int(x) for x in res[2]], int(res[3]), int(res[4]), int(res[5]),
            [[[nm, h(seq)] for (nm, seq) in pair] for pair in res[6]],
            {rn: entry(e) for rn, e in res[7].items()}, {rn: entry(e) for rn, e in res[8].items()}]


def _scan(tables, key, reads):
    """-> (PackedReads oriented as tg_tel.py:265 does, d_cnt_p, d_cnt_q)."""
    from telogator2_b200 import tg_kmer
    T = tables[key]
    KL, CANON = T['KMER_LIST'], T['CANONICAL_STRINGS']
    KLR, CR = [tg_kmer.RC(k) for k in KL], [tg_kmer.RC(k) for k in CANON]
    pr = tg_kmer.PackedReads(reads)
    bc = pr.basecount(CANON, CR).cpu().numpy()
    pr.orient(bc[:, 0] > bc[:, 1])
    d_p, d_q = pr.density_counts(KL, KLR, 100)
    return pr, d_p, d_q


def test_all_bundled_reads_against_reference(reads_all, kmer_tables, boundary_golden):
    from conftest import regions_digest, sha1_of
    from telogator2_b200 import tg_boundary
    P = boundary_golden['params']
    W = P['window']
    total = 0
    for key, lst in reads_all.items():
        reads = [r[1] for r in lst]
        pr, d_p, d_q = _scan(kmer_tables, key, reads)
        G = boundary_golden['sets'][key]
        for pq in 'qp':
            got = tg_boundary.terminating_tl_from_counts(d_p, d_q, pr.d_base_off, pr.lens, W, P['pthresh'], P['min_tel_score'], pq)
            for i, g in enumerate(G):
                want = g['gtt_' + pq]
                inter = None if got[i, 2] < 0 else [int(got[i, 2]), int(got[i, 3])]
                assert [int(got[i, 0]), int(got[i, 1]), inter] == want, (key, pq, lst[i][0])
        # the signal and the region list themselves, read by read
        bb = tg_boundary.BoundaryBatch(np.maximum(pr.lens.astype(np.int64) - W, 0), W)
        bb.power_from_counts(d_p, d_q, pr.d_base_off)
        bb.smooth()
        bb.regions(P['pthresh'])
        for i, g in enumerate(G):
            power = bb.signal_host(i)
            assert len(power) == g['n_power'] and sha1_of(power) == g['power_sha1'], (key, lst[i][0])
            regs = bb.regions_host(i)
            assert len(regs) == g['n_reg'] and regions_digest(regs) == g['regions_sha1'], (key, lst[i][0])
            total += 1
    assert total == 296


def test_get_telomere_regions_synthetic(boundary_golden):
    from telogator2_b200 import tg_kmer
    P = boundary_golden['params']
    cls = {None: 0, 'p': 1, 'q': 2}
    for c in boundary_golden['synthetic']:
        dp, dq = (np.array(c['cnt_p']) / P['window']).tolist(), (np.array(c['cnt_q']) / P['window']).tolist()
        z = [0.0] * c['n']
        if 'error' in c:
            with pytest.raises(IndexError):
                tg_kmer.get_telomere_regions(dp, z, dq, z, P['window'], P['pthresh'])
            continue
        (power, regs) = tg_kmer.get_telomere_regions(dp, z, dq, z, P['window'], P['pthresh'])
        assert np.array_equal(np.asarray(power), np.array(c['power'])), (c['n'], c['kind'])
        assert [[r[0], r[1], cls[r[2]]] for r in regs] == c['regions'], (c['n'], c['kind'])
    # too short: tg_kmer.py:111-112
    assert tg_kmer.get_telomere_regions([0.1] * 50, [0.0] * 50, [0.2] * 50, [0.0] * 50, 100, 0.5) == ([], [[0, 0, None]])
    # smoothing=False keeps the raw signal
    rng = np.random.default_rng(3)
    p, q = rng.integers(0, 101, 700) / 100, rng.integers(0, 101, 700) / 100
    (power, regs) = tg_kmer.get_telomere_regions(p.tolist(), [0.0] * 700, q.tolist(), [0.0] * 700, 100, 0.5, smoothing=False)
    from oracle import boundary as B
    assert np.array_equal(power, (p - q)[:-100]) and regs == B.regions_of((p - q)[:-100], 0.5)


@pytest.mark.parametrize('seed', [1, 2, 3])
def test_batch_against_oracle_bit_for_bit(seed):
    """Seeded count planes of ragged lengths (incl. the sort-in-HBM path: coeff[-5] longer than 4096) vs oracle/boundary.py."""
    import torch
    from oracle import boundary as B
    from telogator2_b200 import tg_boundary
    rng = np.random.default_rng(seed)
    W = 100
    ns = [0, 1, 99, 100, 101, 111, 112, 223, 224, 1000, 1001, 4097, 18423, 65537] + ([140001] if seed == 1 else [30011])
    ns += rng.integers(100, 9000, 40).tolist()
    cps, cqs = [], []
    for n in ns:
        L = n + W
        t = np.arange(n)
        kind = rng.integers(0, 3)
        if kind == 0:
            cp, cq = rng.integers(0, 101, n), rng.integers(0, 101, n)
        elif kind == 1:                                              # telomere-like: q-rich tail, a p-rich island, noise
            cp = np.clip(np.where((t > n // 5) & (t < n // 4), 80, 2) + rng.integers(-2, 9, n), 0, 100)
            cq = np.clip(np.where(t > (2 * n) // 3, 92, 1) + rng.integers(-1, 8, n), 0, 100)
        else:
            cp, cq = np.full(n, 10), np.full(n, 10)
            if n:
                cq[rng.integers(0, n)] = 60
        cps.append(cp.astype(np.uint8)); cqs.append(cq.astype(np.uint8))
    # lay the planes out like PackedReads does (read r at base_off[r], padded to 128)
    lens = np.array(ns, dtype=np.int64) + W
    lens[0] = 37                                                      # a read shorter than the window: no density at all
    ns[0] = 0
    padded = (lens + 128) // 128 * 128
    base_off = np.zeros(len(ns) + 1, dtype=np.int64)
    base_off[1:] = np.cumsum(padded)
    hp = np.zeros(base_off[-1], dtype=np.uint8); hq = np.zeros(base_off[-1], dtype=np.uint8)
    for r, n in enumerate(ns):
        hp[base_off[r]:base_off[r] + n] = cps[r][:n]
        hq[base_off[r]:base_off[r] + n] = cqs[r][:n]
    d_p, d_q, d_off = torch.from_numpy(hp).cuda(), torch.from_numpy(hq).cuda(), torch.from_numpy(base_off).cuda()
    keep = [r for r, n in enumerate(ns) if n != 100]                  # n == window raises (checked below)
    sub = lambda a: [a[r] for r in keep]                              # noqa: E731
    # (select reads by rebuilding the offsets: terminating_tl_from_counts takes consecutive reads)
    for budget in (None, 3_000_000):                                  # one group, and several groups
        for pq in 'qp':
            want = []
            for r in keep:
                dp, dq = cps[r][:ns[r]].astype(np.float64) / W, cqs[r][:ns[r]].astype(np.float64) / W
                (tl, edge, inter) = B.get_terminating_tl(dp, dq, pq, W, 0.5, 100)
                want.append([tl, edge] + ([-1, -1] if inter is None else list(inter)))
            k_off = np.zeros(len(keep) + 1, dtype=np.int64)
            k_off[1:] = np.cumsum(padded[keep])
            kp = np.concatenate([hp[base_off[r]:base_off[r + 1]] for r in keep])
            kq = np.concatenate([hq[base_off[r]:base_off[r + 1]] for r in keep])
            got = tg_boundary.terminating_tl_from_counts(torch.from_numpy(kp).cuda(), torch.from_numpy(kq).cuda(),
                                                         torch.from_numpy(k_off).cuda(), lens[keep], W, 0.5, 100, pq, budget_bytes=budget)
            assert got.tolist() == want, (budget, pq)
    # signals and regions, read by read, exact
    bb = tg_boundary.BoundaryBatch(np.array(sub(ns)), W)
    k_off = np.zeros(len(keep) + 1, dtype=np.int64)
    k_off[1:] = np.cumsum(padded[keep])
    bb.power_from_counts(torch.from_numpy(np.concatenate([hp[base_off[r]:base_off[r + 1]] for r in keep])).cuda(),
                         torch.from_numpy(np.concatenate([hq[base_off[r]:base_off[r + 1]] for r in keep])).cuda(),
                         torch.from_numpy(k_off).cuda())
    bb.smooth()
    bb.regions(0.5)
    n_smoothed = 0
    for i, r in enumerate(keep):
        dp, dq = cps[r][:ns[r]].astype(np.float64) / W, cqs[r][:ns[r]].astype(np.float64) / W
        (power, regs) = B.get_telomere_regions(dp, dp, dq, dq, W, 0.5)
        assert regs == bb.regions_host(i), ns[r]
        if ns[r] >= W:
            assert np.array_equal(bb.signal_host(i), power), ns[r]
            n_smoothed += int(bb.d_smoothed[i].item())
    assert n_smoothed > 20
    # a read whose density is exactly one window long: the reference raises IndexError at p_vs_q_power[0]
    bb = tg_boundary.BoundaryBatch(np.array([100]), W)
    bb.power_from_host([np.zeros(100)])
    bb.smooth()
    with pytest.raises(IndexError):
        bb.regions(0.5)


def test_wavelet_smooth_wrapper():
    from oracle import boundary as B
    from telogator2_b200 import tg_kmer
    rng = np.random.default_rng(9)
    for n in [5, 111, 112, 113, 1000, 1001, 8191, 20000]:
        x = rng.standard_normal(n) * 0.3 + np.sin(np.arange(n) / 50.0)
        want = B.wavelet_smooth(x)
        got = tg_kmer.wavelet_smooth(x)
        if want is x:
            assert got is x
        else:
            assert np.array_equal(got, want), n
    x = np.zeros(3000)
    assert tg_kmer.wavelet_smooth(x) is x                            # sigma below fail_sigma
    assert tg_kmer.mad(np.array([1.0, 2.0, 4.0, 9.0])) == B.mad(np.array([1.0, 2.0, 4.0, 9.0]))


def test_get_terminating_tl_wrapper(reads_all, kmer_tables, boundary_golden):
    from telogator2_b200 import tg_kmer, tg_tel
    P = boundary_golden['params']
    for key in ('human_hifi', 'maize_hifi'):
        T = kmer_tables[key]
        KL, CANON = T['KMER_LIST'], T['CANONICAL_STRINGS']
        KLR, CR = [tg_kmer.RC(k) for k in KL], [tg_kmer.RC(k) for k in CANON]
        for (name, rdat), g in list(zip(reads_all[key], boundary_golden['sets'][key]))[:12]:
            ori = tg_kmer.RC(rdat) if tg_kmer.get_telomere_base_count(rdat, CANON) > tg_kmer.get_telomere_base_count(rdat, CR) else rdat
            for pq in 'qp':
                (tl, edge, inter) = tg_tel.get_terminating_tl(ori, pq, [KL, KLR, P['window'], P['pthresh'], P['min_tel_score']])
                assert [tl, edge, None if inter is None else list(inter)] == g['gtt_' + pq], (name, pq)
    assert tg_tel.get_terminating_tl('ACGT' * 10, 'q', [KL, KLR, 100, 0.5, 100]) == (0, 0, None)


@pytest.mark.parametrize('tkey,filters', gtt_cases())
def test_batch_boundary_with_device_boundary_calling_matches_reference_driver(reads_all, kmer_tables, tkey, filters):
    """tg_tel.get_tel_repeat_comp_parallel, default path (boundary calling on the device), vs the reference's own
    get_tel_repeat_comp_parallel + get_terminating_tl on all bundled reads (tests/golden/make_golden_tel_gtt.py)."""
    from telogator2_b200 import tg_kmer, tg_tel
    with gzip.open(os.path.join(HERE, 'golden', 'tel_batch_gtt_golden.json.gz'), 'rt') as f:
        want = json.load(f)[tkey + '|' + ','.join(map(str, filters))]
    reads = [(name + ' extra', seq, None) for name, seq in reads_all[tkey]]
    params = gtt_params_of(kmer_tables[tkey], filters, tg_kmer.RC)
    got = tg_tel.get_tel_repeat_comp_parallel(reads, params, max_workers=7)
    assert tg_tel.LAST_STATS['boundary'] == 'device'
    assert compact_tel_result(got) == want
    # and the callback path (what the telomere-signal plots use) with the numpy oracle as the callback gives the same lists
    from oracle import boundary as B

    def gtt(rdat, pq, gtt_params, telplot_dat=None):
        [KL, KLR, W, thr, min_score] = gtt_params
        return B.get_terminating_tl(np.asarray(tg_kmer.get_telomere_kmer_density(rdat, KL, W)[0]),
                                    np.asarray(tg_kmer.get_telomere_kmer_density(rdat, KLR, W)[0]), pq, W, thr, min_score)
    got_cb = tg_tel.get_tel_repeat_comp_parallel(reads, params, get_terminating_tl=gtt)
    assert tg_tel.LAST_STATS['boundary'] == 'callback'
    assert compact_tel_result(got_cb) == want
