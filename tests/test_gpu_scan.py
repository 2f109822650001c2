"""GPU parity: k-mer scan (tg_kmer) through the C ABI vs golden vectors from the reference and the oracle."""
import numpy as np
import pytest

import oracle
from conftest import hits_digest, hits_to_array, rc_py, scan_cases, sha1_of

pytestmark = pytest.mark.gpu


@pytest.fixture(scope='module')
def tgk():
    import torch
    assert torch.cuda.is_available()
    from telogator2_b200 import tg_kmer
    return tg_kmer


def test_batch_scan_matches_reference_golden(tgk, kmer_tables, golden_reads, scan_golden):
    """Phase 1 of the batch boundary: base counts, orientation, densities for every golden read."""
    by_table = {}
    for tag, T, rdat in scan_cases(kmer_tables, golden_reads):
        by_table.setdefault(tag.split('|')[1], []).append((tag, rdat))
    n_checked = 0
    for tkey, items in by_table.items():
        T = kmer_tables[tkey]
        KL, CANON = T['KMER_LIST'], T['CANONICAL_STRINGS']
        KL_REV, CANON_REV = [rc_py(k) for k in KL], [rc_py(k) for k in CANON]
        reads = [r for _, r in items]
        pr = tgk.PackedReads(reads)
        bc = pr.basecount(CANON, CANON_REV).cpu().numpy()
        for q, (tag, _r) in enumerate(items):
            assert bc[q].tolist() == scan_golden[tag + '|bc'].tolist(), tag
        flip = bc[:, 0] > bc[:, 1]
        pr.orient(flip)
        ca, cb = pr.density_counts(KL, KL_REV, 100)
        ca, cb = pr.split_counts(ca.cpu().numpy(), 100), pr.split_counts(cb.cpu().numpy(), 100)
        for q, (tag, _r) in enumerate(items):
            assert int(flip[q]) == int(scan_golden[tag + '|flip'][0]), tag
            gp, gq = scan_golden[tag + '|dens_p'], scan_golden[tag + '|dens_q']
            assert ca[q].shape == gp.shape, tag
            bad = np.nonzero(ca[q] != gp)[0]
            assert len(bad) == 0, f'{tag} dens_p differs at {bad[:8]}: gpu {ca[q][bad[:8]]} ref {gp[bad[:8]]}'
            bad = np.nonzero(cb[q] != gq)[0]
            assert len(bad) == 0, f'{tag} dens_q differs at {bad[:8]}: gpu {cb[q][bad[:8]]} ref {gq[bad[:8]]}'
            n_checked += 1
        # hits on the oriented reads: suffix and prefix slices in one batch
        ISSUB = T['KMER_ISSUBSTRING']
        lens = [len(r) for r in reads]
        idx = np.arange(len(reads))
        out_sfx = pr.hits(idx, [max(l - 6000, 0) for l in lens], lens, KL_REV, ISSUB)
        out_fp = pr.hits(idx, [0] * len(reads), [min(l, 20000) for l in lens], KL, ISSUB)
        out_fq = pr.hits(idx, [0] * len(reads), [min(l, 20000) for l in lens], KL_REV, ISSUB)
        for q, (tag, _r) in enumerate(items):
            for nm, got in (('suffix_q', out_sfx[q]), ('full_p', out_fp[q]), ('full_q', out_fq[q])):
                g = hits_to_array(got)
                want = scan_golden[tag + '|hits_' + nm]
                assert g.shape == want.shape and np.array_equal(g, want), f'{tag} hits_{nm}: gpu {g.shape} ref {want.shape}'
    assert n_checked > 60


def test_per_read_functions_keep_reference_types(tgk, kmer_tables, golden_reads, scan_golden):
    T = kmer_tables['human_hifi']
    KL, ISSUB, CANON = T['KMER_LIST'], T['KMER_ISSUBSTRING'], T['CANONICAL_STRINGS']
    rdat = golden_reads['human_hifi'][0][1]
    tag = 'human_hifi|human_hifi|0'
    bc = tgk.get_telomere_base_count(rdat, CANON)
    assert isinstance(bc, int) and bc == int(scan_golden[tag + '|bc'][0])
    ori = rc_py(rdat) if int(scan_golden[tag + '|flip'][0]) else rdat
    e0, e1 = tgk.get_telomere_kmer_density(ori, KL, 100)
    assert isinstance(e0, list) and isinstance(e1, list) and len(e0) == len(e1) == len(ori) - 100
    assert e0 == (scan_golden[tag + '|dens_p'].astype(np.float64) / 100).tolist() and all(v == 0.0 for v in e1)
    assert tgk.get_telomere_kmer_density('ACGT' * 10, KL, 100) == ([], [])
    hits = tgk.get_nonoverlapping_kmer_hits(ori[-6000:], [rc_py(k) for k in KL], ISSUB)
    assert isinstance(hits, list) and len(hits) == len(KL) and all(isinstance(s, list) and len(s) == 2 for h in hits for s in h)
    assert np.array_equal(hits_to_array(hits), scan_golden[tag + '|hits_suffix_q'])
    assert tgk.get_nonoverlapping_kmer_hits('', KL, ISSUB) == [[] for _ in KL]
    assert tgk.get_telomere_base_count('', CANON) == 0


def test_random_and_adversarial_vs_oracle(tgk, kmer_tables):
    rng = np.random.default_rng(11)
    T = kmer_tables['human_ont']
    KL, ISSUB = T['KMER_LIST'], T['KMER_ISSUBSTRING']
    reads = []
    for n in (1, 2, 99, 100, 101, 127, 128, 129, 256, 8191, 8192, 8193, 8292, 16500, 32576, 32577, 40000, 65153):
        reads.append(''.join(rng.choice(list('ACGT'), n, p=[.25, .45, .1, .2])))
    reads.append('C' * 20000)                                    # longest possible greedy chain across tiles
    reads.append('C' * 70001)
    reads.append('CCCTAA' * 5400 + 'CCCCC' + 'CCCCTAACCCTAACCCCTAA' * 30 + 'CCCTCC' * 7 + 'CCCTAA' * 6000)   # 32 k tile borders inside repeats
    reads.append('ACGTTGCA' * 4060 + 'C' * 200 + 'ACGTTGCA' * 4000)         # a chain of candidates across the 32576 border
    reads.append(('CCCTCC' * 3 + 'CCCTAA' * 2) * 1500)
    reads.append('CCCTAA' * 1365 + 'CCCCCC' * 10 + 'CCCTAA' * 1400)   # tile boundary inside repeats
    reads.append(''.join(rng.choice(list('ACGTN'), 9000, p=[.2, .5, .1, .15, .05])))
    reads.append('acgtCCCTAACCCTAAxyzCCCTAA' * 40)              # lower case / junk never matches
    pr = tgk.PackedReads(reads)
    for lists in ((KL, [rc_py(k) for k in KL]), (['CCCCCC', 'CCC'], ['GGGTTA'])):
        bc = pr.basecount(*lists).cpu().numpy()
        for w in (100, 7, 255):
            ca, cb = pr.density_counts(lists[0], lists[1], w)
            ca, cb = pr.split_counts(ca.cpu().numpy(), w), pr.split_counts(cb.cpu().numpy(), w)
            for q, r in enumerate(reads):
                assert np.array_equal(ca[q].astype(np.int32), oracle.density_counts(r, lists[0], w)), (q, w)
                assert np.array_equal(cb[q].astype(np.int32), oracle.density_counts(r, lists[1], w)), (q, w)
        for q, r in enumerate(reads):
            assert bc[q].tolist() == [oracle.base_count(r, lists[0]), oracle.base_count(r, lists[1])], q
    short = [r for r in reads if len(r) <= 20000]
    pr = tgk.PackedReads(short)
    lens = [len(r) for r in short]
    got = pr.hits(np.arange(len(short)), [0] * len(short), lens, KL, ISSUB)
    for q, r in enumerate(short):
        assert np.array_equal(hits_to_array(got[q]), hits_to_array(oracle.nonoverlapping_hits(r, KL, ISSUB))), q
    # unaligned slices of one read
    r = reads[22]
    sl = [(3, 4100), (2047, 2049), (100, 100), (5, 37), (4090, 8200)]
    pr = tgk.PackedReads([r])
    got = pr.hits([0] * len(sl), [a for a, _ in sl], [b for _, b in sl], KL, ISSUB)
    for q, (a, b) in enumerate(sl):
        assert np.array_equal(hits_to_array(got[q]), hits_to_array(oracle.nonoverlapping_hits(r[a:b], KL, ISSUB))), (a, b)


def test_orientation_rejects_non_acgtn_like_reference(tgk):
    pr = tgk.PackedReads(['CCCTAACCCTAAxCCCTAA'])
    with pytest.raises(KeyError):
        pr.orient(np.array([1]))


def test_scale_properties(tgk, kmer_tables, golden_reads):
    """~100 Mbases replicated set: density of a replicated read equals the density of the original."""
    T = kmer_tables['human_hifi']
    KL = T['KMER_LIST']
    KL_REV = [rc_py(k) for k in KL]
    base = [r for _, r in golden_reads['human_hifi']]
    reads = base * 400
    pr = tgk.PackedReads(reads)
    ca, cb = pr.density_counts(KL, KL_REV, 100)
    ca = pr.split_counts(ca.cpu().numpy(), 100)
    for q in range(len(base), len(reads), 37):
        assert np.array_equal(ca[q], ca[q % len(base)])
    bc = pr.basecount(KL, KL_REV).cpu().numpy()
    assert np.array_equal(bc[:len(base)], bc[-len(base):])


def test_bench_workload_sample_against_oracle(tgk, kmer_tables, reads_all):
    """The bench's scan workload at 60 Mbases (bundled HiFi + ONT reads replicated with seeded edits: every read different, tile
    and read boundaries at every alignment): base counts, orientation, both densities, the hits of the last 6000 bases and
    the boundary calling of a sample of reads, each against the CPU oracle on that read alone."""
    from oracle import boundary as B
    from telogator2_b200 import tg_boundary
    rng = np.random.default_rng(99)
    base = [r for _, r in reads_all['human_hifi']] + [r for _, r in reads_all['human_ont']]
    acgt = np.frombuffer(b'ACGTN', dtype=np.uint8)
    reads, tot, i = [], 0, 0
    while tot < 60_000_000:
        a = np.frombuffer(base[i % len(base)].encode(), dtype=np.uint8).copy()
        k = max(len(a) // 150, 1)
        a[rng.integers(len(a), size=k)] = acgt[rng.integers(4 + (i % 7 == 0), size=k)]       # (every 7th read also gets Ns)
        reads.append(a.tobytes().decode())
        tot += len(a)
        i += 1
    T = kmer_tables['human_hifi']
    KL, ISSUB, CANON = T['KMER_LIST'], T['KMER_ISSUBSTRING'], T['CANONICAL_STRINGS']
    KLR, CR = [rc_py(k) for k in KL], [rc_py(k) for k in CANON]
    pr = tgk.PackedReads(reads)
    bc = pr.basecount(CANON, CR).cpu().numpy()
    flip = bc[:, 0] > bc[:, 1]
    pr.orient(flip)
    d_p, d_q = pr.density_counts(KL, KLR, 100)
    term = tg_boundary.terminating_tl_from_counts(d_p, d_q, pr.d_base_off, pr.lens, 100, 0.5, 100, 'q')
    cp, cq = pr.split_counts(d_p.cpu().numpy(), 100), pr.split_counts(d_q.cpu().numpy(), 100)
    sample = np.sort(rng.choice(len(reads), 150, replace=False))
    lens = pr.lens.astype(np.int64)
    hits = pr.hits(sample, np.maximum(lens[sample] - 6000, 0), lens[sample], KLR, ISSUB)
    for q, r in enumerate(sample.tolist()):
        fl, ori, c0, c1, obc = oracle.scan_read(reads[r], KL, KLR, CANON, CR, 100)
        assert bc[r].tolist() == obc.tolist() and bool(flip[r]) == fl, r
        assert np.array_equal(cp[r], c0) and np.array_equal(cq[r], c1), r
        assert hits[q] == oracle.nonoverlapping_hits(ori[-6000:], KLR, ISSUB), r
        (tl, edge, inter) = B.get_terminating_tl(c0.astype(np.float64) / 100, c1.astype(np.float64) / 100, 'q', 100, 0.5, 100)
        assert term[r].tolist() == [tl, edge] + ([-1, -1] if inter is None else list(inter)), r


def test_limits_fail_loudly(tgk, kmer_tables):
    T = kmer_tables['human_hifi']
    KL = T['KMER_LIST']
    pr = tgk.PackedReads(['CCCTAA' * 100])
    with pytest.raises(Exception):
        pr.density_counts(KL, KL, 256)                         # uint8 counts: window <= 255
    with pytest.raises(NotImplementedError):
        tgk.get_telomere_kmer_density('CCCTAA' * 100, KL, 100, smoothing=True)
    with pytest.raises(NotImplementedError):
        tgk.kmer_table(['ACGT' + format(i, '06b').replace('0', 'A').replace('1', 'C') for i in range(65)])
    with pytest.raises(Exception):
        tgk.kmer_table(['ACGT' * 9])                           # 36 > 32 bases
    with pytest.raises(Exception):
        tgk.kmer_table(['ACNT'])


def test_window_and_short_reads(tgk, kmer_tables):
    T = kmer_tables['maize_hifi']
    KL = T['KMER_LIST']
    reads = ['', 'C', 'CCCTAAA', 'CCCTAAA' * 14 + 'CC', 'CCCTAAA' * 15, 'CCCTAAA' * 40]
    pr = tgk.PackedReads(reads)
    for w in (1, 3, 100, 255):
        ca, cb = pr.density_counts(KL, [tgk.RC(k) for k in KL], w)
        ca = pr.split_counts(ca.cpu().numpy(), w)
        for q, r in enumerate(reads):
            assert np.array_equal(ca[q].astype(np.int32), oracle.density_counts(r, KL, w)), (q, w)


def test_all_bundled_reads_match_reference(tgk, kmer_tables, reads_all, scan_all_golden):
    """BASELINE configs 1-3 whole (83 ONT + 200 HiFi + 13 maize reads): base counts, orientation, both density tracks and the
    hit spans of the last 6000 bases, against what the reference's own tg_kmer functions computed (make_golden_all.py)."""
    n = 0
    for key, rs in reads_all.items():
        T = kmer_tables[key]
        KL, ISSUB, CANON = T['KMER_LIST'], T['KMER_ISSUBSTRING'], T['CANONICAL_STRINGS']
        KLR, CR = [rc_py(k) for k in KL], [rc_py(k) for k in CANON]
        reads = [r for _, r in rs]
        pr = tgk.PackedReads(reads)
        bc = pr.basecount(CANON, CR).cpu().numpy()
        flip = bc[:, 0] > bc[:, 1]
        pr.orient(flip)
        ca, cb = pr.density_counts(KL, KLR, 100)
        ca, cb = pr.split_counts(ca.cpu().numpy(), 100), pr.split_counts(cb.cpu().numpy(), 100)
        lens = np.array([len(r) for r in reads])
        hits = pr.hits(np.arange(len(reads)), np.maximum(lens - 6000, 0), lens, KLR, ISSUB)
        for q, ((name, _r), g) in enumerate(zip(rs, scan_all_golden[key])):
            assert bc[q].tolist() == g['bc'] and bool(flip[q]) == g['flip'], name
            assert int(ca[q].astype(np.int64).sum()) == g['sum_p'] and int(cb[q].astype(np.int64).sum()) == g['sum_q'], name
            assert sha1_of(ca[q]) == g['dens_p'] and sha1_of(cb[q]) == g['dens_q'], name
            assert hits_digest(hits[q]) == (g['hits'], g['n_spans']), name
            n += 1
    assert n == 296
