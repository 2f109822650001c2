#!/usr/bin/env python
"""Generate the golden fixtures in tests/golden/ from the REFERENCE's own code.

Run in the build container only (needs /root/reference; the GPU box never runs this):

    python tests/golden/make_golden.py

What is produced (all small, committed):
  kmer_tables.json   read_kmer_tsv() outputs + raw TSV text for the bundled k-mer tables
  reads.json.gz      a subset of the bundled reads + adversarial synthetic strings
  scan_golden.npz    per (case) reference outputs of get_telomere_base_count,
                     get_telomere_kmer_density, get_nonoverlapping_kmer_hits, RC,
                     and the scan part of gtrc_parallel_job (orientation)
  rng_golden.npz     CPython random.seed()/random.sample() permutations (the reference RNG)
  colorvecs.json.gz  colour vectors (TVR letter strings) of the bundled reads: DP inputs
  dp_golden.npz      distance-matrix fixtures from the CPU oracle (NW score is PARITY UNPINNED:
                     Biopython is not installable here, see oracle/tg_oracle.c)

The reference modules are imported unmodified; third-party modules that are absent in this
container (pywt, matplotlib, Bio, pysam) are replaced by inert stubs -- none of the functions
called here touches them (wavelet_smooth is only reached from get_telomere_regions, which is
replaced by an identity smoother when colour vectors are derived; that only affects where the
TVR suffix starts, i.e. which realistic strings become DP inputs, not any golden output of the
scan functions).
"""
import gzip
import json
import os
import random
import sys
import types
from unittest import mock

import numpy as np

REF = '/root/reference'
HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, REF)
sys.path.insert(1, ROOT)
for name in ['pywt', 'matplotlib', 'matplotlib.pyplot', 'matplotlib.colors', 'matplotlib.cm', 'matplotlib.lines',
             'matplotlib.patches', 'matplotlib.collections', 'Bio', 'Bio.Align', 'pysam']:
    sys.modules[name] = mock.MagicMock()

from source import tg_kmer, tg_tel, tg_util  # noqa: E402  (the reference)
import oracle                                   # noqa: E402

tg_kmer.wavelet_smooth = lambda x, *a, **k: x   # identity smoother (pywt absent); see docstring

TEL_WINDOW_SIZE = 100          # telogator2.py:22


def read_fa(fn):
    out, name, seq = [], None, []
    for line in gzip.open(fn, 'rt'):
        if line[0] == '>':
            if name:
                out.append((name, ''.join(seq)))
            name, seq = line[1:].strip(), []
        else:
            seq.append(line.strip())
    if name:
        out.append((name, ''.join(seq)))
    return out


def flatten_hits(hits):
    k, s, e = [], [], []
    for ki, spans in enumerate(hits):
        for sp in spans:
            assert isinstance(sp, list) and len(sp) == 2
            k.append(ki), s.append(sp[0]), e.append(sp[1])
    return np.array(k, np.int32), np.array(s, np.int32), np.array(e, np.int32)


def main():
    rng = np.random.default_rng(20240607)
    tables = {}
    specs = [('human_ont', 'source/resources/kmers.tsv', 'ont'),
             ('human_hifi', 'source/resources/kmers.tsv', 'hifi'),
             ('maize_hifi', 'source/resources/non-human/kmers_maize.tsv', 'hifi'),
             ('celegans_ont', 'source/resources/non-human/kmers_celegans.tsv', 'ont')]
    for key, path, rtype in specs:
        fn = os.path.join(REF, path)
        (meta, issub, canon) = tg_kmer.read_kmer_tsv(fn, rtype)
        tables[key] = {'tsv_text': open(fn).read(), 'read_type': rtype,
                       'KMER_LIST': meta[0], 'KMER_COLORS': meta[1], 'KMER_LETTER': meta[2],
                       'KMER_FLAGS': [list(f) for f in meta[3]], 'KMER_ISSUBSTRING': issub,
                       'CANONICAL_STRINGS': canon, 'canonical_letter': tg_kmer.get_canonical_letter(fn, rtype)}
    json.dump(tables, open(os.path.join(HERE, 'kmer_tables.json'), 'w'), indent=0)

    # ---------------- reads: subset of bundled + adversarial ----------------
    sets = {'human_ont': read_fa(f'{REF}/test_data/hg002-ont-1p.fa.gz'),
            'human_hifi': read_fa(f'{REF}/test_data/hg002-telreads_pacbio.sub.fa.gz'),
            'maize_hifi': read_fa(f'{REF}/test_data/ZMMo17-hifi-7p8p.fa.gz')}
    reads = {}
    for key, rs in sets.items():
        order = sorted(range(len(rs)), key=lambda i: len(rs[i][1]))
        pick = order[:6] + order[len(order) // 2: len(order) // 2 + 3] + order[-1:]
        if key == 'maize_hifi':
            pick = order[:5]
        reads[key] = [[rs[i][0], rs[i][1]] for i in sorted(set(pick))]
    syn = []
    syn.append(['syn_empty', ''])
    syn.append(['syn_short', 'CCCTAACCCTAA'])
    syn.append(['syn_len100', 'CCCTAA' * 16 + 'CCCT'])
    syn.append(['syn_len101', 'CCCTAA' * 16 + 'CCCTA'])
    syn.append(['syn_polyC', 'C' * 777])
    syn.append(['syn_polyC_tail', 'ACGT' * 10 + 'C' * 403 + 'TTAGGG' * 30])
    syn.append(['syn_cccTcc', 'CCCTCCCTCCCTCCCTCC' * 20 + 'CCCTCC' * 7 + 'CCTCCC' * 9])
    syn.append(['syn_hor', ('CCCCTACCCTAACCCCTA' * 3 + 'CCCTAA' * 2 + 'CCCCTAACCCTAACCCCTAA' * 3) * 8])
    syn.append(['syn_hor_overlap', 'CCCCTACCCTAACCCCTACCCTAACCCCTACCCTAACCCCTA' * 5 + 'CCCCTAACCCTAACCCCTAACCCTAACCCCTAACCCTAACCCCTAA' * 5])
    syn.append(['syn_N', 'CCCTAA' * 30 + 'N' + 'CCCTAA' * 5 + 'CCCNAACCCTAN' * 6 + 'TTAGGG' * 25 + 'NNNN' + 'TTAGGG' * 10])
    syn.append(['syn_q_telomere', ''.join(rng.choice(list('ACGT'), 1500)) + 'TTAGGG' * 40 + 'TTGGGG' * 3 + 'TCAGGG' * 5 + 'TTAGGG' * 300])
    syn.append(['syn_p_telomere', 'CCCTAA' * 250 + 'CCCTGA' * 4 + 'CCCCAA' * 3 + 'CCCTAA' * 20 + ''.join(rng.choice(list('ACGT'), 1800))])
    syn.append(['syn_random', ''.join(rng.choice(list('ACGT'), 5000))])
    syn.append(['syn_random_C_rich', ''.join(rng.choice(list('ACGT'), 6000, p=[.2, .6, .05, .15]))])
    # noisy telomere: repeats with random edits (mimics ONT errors) -> many overlapping k-mer hits
    tel = list('CCCTAA' * 900)
    for pos in sorted(rng.choice(len(tel), 500, replace=False), reverse=True):
        r = rng.integers(3)
        if r == 0:
            tel[pos] = rng.choice(list('ACGT'))
        elif r == 1:
            del tel[pos]
        else:
            tel.insert(pos, rng.choice(list('ACGT')))
    syn.append(['syn_noisy_p', ''.join(tel) + ''.join(rng.choice(list('ACGT'), 1200))])
    reads['synthetic'] = syn
    with gzip.open(os.path.join(HERE, 'reads.json.gz'), 'wt') as f:
        json.dump(reads, f)

    # ---------------- scan golden ----------------
    gold = {}
    colorvecs = {}
    for rkey, rlist in reads.items():
        tkeys = [rkey] if rkey != 'synthetic' else ['human_ont', 'human_hifi', 'maize_hifi']
        for tkey in tkeys:
            T = tables[tkey]
            KL, ISSUB, CANON = T['KMER_LIST'], T['KMER_ISSUBSTRING'], T['CANONICAL_STRINGS']
            KL_REV = [tg_util.RC(k) for k in KL]
            CANON_REV = [tg_util.RC(k) for k in CANON]
            for ri, (rname, rdat) in enumerate(rlist):
                tag = f'{rkey}|{tkey}|{ri}'
                bc_f = tg_kmer.get_telomere_base_count(rdat, CANON, mode=T['read_type'])
                bc_r = tg_kmer.get_telomere_base_count(rdat, CANON_REV, mode=T['read_type'])
                gold[tag + '|bc'] = np.array([bc_f, bc_r], np.int32)
                ori = tg_util.RC(rdat) if bc_f > bc_r else rdat          # tg_tel.py:265-266
                gold[tag + '|flip'] = np.array([int(bc_f > bc_r)], np.int8)
                for sname, klist in (('p', KL), ('q', KL_REV)):           # tg_tel.py:160-161
                    (e0, e1) = tg_kmer.get_telomere_kmer_density(ori, klist, TEL_WINDOW_SIZE)
                    assert isinstance(e0, list) and len(e0) == len(e1) == max(len(ori) - TEL_WINDOW_SIZE, 0)
                    assert all(v == 0.0 for v in e1)
                    cnt = np.rint(np.array(e0, dtype=np.float64) * TEL_WINDOW_SIZE).astype(np.uint8)
                    assert np.array_equal(cnt.astype(np.float64) / TEL_WINDOW_SIZE, np.array(e0, dtype=np.float64))
                    gold[tag + '|dens_' + sname] = cnt
                # hits on a TVR+tel style suffix and on the whole read, both k-mer lists (tg_tel.py:287-290,310)
                suffix = ori[-6000:]
                for sname, s, klist in (('suffix_q', suffix, KL_REV), ('full_p', ori[:20000], KL), ('full_q', ori[:20000], KL_REV)):
                    hits = tg_kmer.get_nonoverlapping_kmer_hits(s, klist, ISSUB)
                    k, st, en = flatten_hits(hits)
                    gold[tag + '|hits_' + sname] = np.stack([k, st, en]) if len(k) else np.zeros((3, 0), np.int32)
                # colour vector of the suffix hits (tg_tvr.py:96-105) -> DP inputs
                if rkey != 'synthetic':
                    gtt = [KL, KL_REV, TEL_WINDOW_SIZE, 0.5, 100]
                    (tl, _nontel, _i) = tg_tel.get_terminating_tl(ori, 'q', gtt)
                    tl = min(tl, len(ori))
                    sub_end = max(len(ori) - tl - 1000, 0)
                    tvrseq = ori[sub_end:] or ori
                    hits = tg_kmer.get_nonoverlapping_kmer_hits(tvrseq, KL_REV, ISSUB)
                    letters = ['A'] * len(tvrseq)
                    for ki in range(len(hits)):
                        for sp in hits[ki]:
                            for j in range(sp[0], sp[1]):
                                letters[j] = T['KMER_LETTER'][ki]
                    colorvecs.setdefault(tkey, []).append(''.join(letters))
    np.savez_compressed(os.path.join(HERE, 'scan_golden.npz'), **gold)

    # colour vectors for ALL bundled reads (DP workloads): strip leading subtel-ish part crudely by
    # keeping the last 1000+TL letters -- realistic letter mix is what matters for DP inputs
    cv_all = {}
    for key, rs in sets.items():
        T = tables[key]
        KL, ISSUB, CANON = T['KMER_LIST'], T['KMER_ISSUBSTRING'], T['CANONICAL_STRINGS']
        KL_REV = [tg_util.RC(k) for k in KL]
        CANON_REV = [tg_util.RC(k) for k in CANON]
        out = []
        for (rname, rdat) in rs:
            bc_f = tg_kmer.get_telomere_base_count(rdat, CANON)
            bc_r = tg_kmer.get_telomere_base_count(rdat, CANON_REV)
            ori = tg_util.RC(rdat) if bc_f > bc_r else rdat
            (tl, nontel, _i) = tg_tel.get_terminating_tl(ori, 'q', [KL, KL_REV, TEL_WINDOW_SIZE, 0.5, 100])
            tl = min(tl, len(ori))
            if tl < 400 or nontel > 100 or len(ori) < tl + 1000:       # default --filt-* (telogator2.py:50-52)
                continue
            tvrseq = ori[max(len(ori) - tl - 1000, 0):]
            hits = oracle.nonoverlapping_hits(tvrseq, KL_REV, ISSUB)    # == reference (pinned below)
            letters = ['A'] * len(tvrseq)
            for ki in range(len(hits)):
                for sp in hits[ki]:
                    for j in range(sp[0], sp[1]):
                        letters[j] = T['KMER_LETTER'][ki]
            cv = ''.join(letters)
            # crude subtel strip: start at the first window of 50 with <= 50% unknown
            start = 0
            for i in range(0, max(len(cv) - 50, 1)):
                if cv[i:i + 50].count('A') <= 25:
                    start = i
                    break
            while start < len(cv) - 1 and cv[start] == 'A':
                start += 1
            out.append(cv[start:])
        cv_all[key] = out
        print(key, 'colour vectors:', len(out), 'median len', int(np.median([len(c) for c in out])))
    with gzip.open(os.path.join(HERE, 'colorvecs.json.gz'), 'wt') as f:
        json.dump({'colorvecs': cv_all, 'canonical_letter': {k: tables[k]['canonical_letter'] for k in cv_all}}, f)

    # ---------------- RNG golden (CPython random == the reference's RNG) ----------------
    rg = {}
    for seed in [0, 1, 12345, 12345 + 3402, 99999999, 2 ** 32 - 1, 2 ** 32, 2 ** 40 + 7, 2 ** 63 + 11]:
        for n in [1, 2, 3, 5, 6, 7, 31, 32, 33, 100, 623, 624, 625, 1000, 3000]:
            random.seed(seed)
            base = ''.join(chr(33 + (i % 90)) for i in range(n))
            idx = {}
            p1 = random.sample(range(n), n)
            p2 = random.sample(range(max(n - 1, 1)), max(n - 1, 1))
            p3 = random.sample(range(n), n)
            rg[f'{seed}|{n}'] = np.concatenate([p1, p2, p3]).astype(np.int32)
            random.seed(seed)
            s1 = tg_util.shuffle_seq(base)
            assert s1 == ''.join(base[i] for i in p1)       # permutation depends on (seed, n) only
            del idx
    np.savez_compressed(os.path.join(HERE, 'rng_golden.npz'), **rg)

    # ---------------- DP golden (oracle; NW score parity unpinned) ----------------
    dp = {}
    P_tvr = oracle.AlignerParams(oracle.scoring_matrix('C', 'tvr'), (True, True))
    P_ms = oracle.AlignerParams(None, (True, False))
    ont = [c[:1000] for c in cv_all['human_ont']][:24]
    hifi = [c[:1000] for c in cv_all['human_hifi']][:24]
    for name, seqs, P, adj, mv, R in (('ont24_tvr_R2', ont, P_tvr, True, True, 2),
                                      ('hifi24_tvr_R2', hifi, P_tvr, True, True, 2),
                                      ('ont12_tvr2000_R3', [c[:2000] for c in cv_all['human_ont']][:12], P_tvr, True, True, 3),
                                      ('ont12_ms_rightfree_R3', [c[:700 + 97 * i] for i, c in enumerate(cv_all['human_ont'][:12])], P_ms, False, False, 3)):
        D, aln, rnd, cells = oracle.dist_matrix_c(seqs, P, adj, mv, R, 12345, want_scores=True)
        # cross-check the C path against the pure-Python restatement that uses CPython's `random`
        det = {}
        for p, (i, j) in enumerate([(0, 1), (0, 2), (3, 7)]):
            pi = [q for q, ij in enumerate(__import__('itertools').combinations(range(len(seqs)), 2)) if ij == (i, j)][0]
            d = oracle.tvr_distance_py(seqs[i], seqs[j], P, adj, mv, R, 12345 + pi, detail=det)
            assert np.float32(d) == D[i, j], (name, i, j, d, D[i, j])
            assert int(det['aln']) == aln[pi]
        dp[name + '|D'] = D
        dp[name + '|aln'] = aln
        dp[name + '|rnd'] = rnd
        dp[name + '|cells'] = np.array([cells])
        print(name, 'early-out frac', float(np.mean(rnd[:, 0] == np.iinfo(np.int32).min)))
    np.savez_compressed(os.path.join(HERE, 'dp_golden.npz'), **dp)
    print('done')


if __name__ == '__main__':
    main()
