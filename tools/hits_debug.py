#!/usr/bin/env python
"""GPU vs oracle for get_nonoverlapping_kmer_hits on the golden reads: prints the first differing spans with context."""
import gzip, json, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import oracle
from telogator2_b200 import tg_kmer
G = os.path.join(ROOT, 'tests', 'golden')
T = json.load(open(os.path.join(G, 'kmer_tables.json')))
R = json.load(gzip.open(os.path.join(G, 'reads.json.gz'), 'rt'))
nbad = 0
for tname in ('human_hifi', 'human_ont', 'maize_hifi'):
    KL, ISS = T[tname]['KMER_LIST'], T[tname]['KMER_ISSUBSTRING']
    KR = [tg_kmer.RC(k) for k in KL]
    reads = [r[1] for r in R[tname]]
    pr = tg_kmer.PackedReads(reads)
    for lst, nm in ((KL, 'p'), (KR, 'q')):
        got = pr.hits(np.arange(len(reads)), [0] * len(reads), [min(len(r), 20000) for r in reads], lst, ISS)
        for i, r in enumerate(reads):
            want = oracle.nonoverlapping_hits(r[:20000], lst, ISS)
            if got[i] == want:
                continue
            nbad += 1
            gs = sorted((s, e, k) for k in range(len(lst)) for s, e in got[i][k])
            ws = sorted((s, e, k) for k in range(len(lst)) for s, e in want[k])
            only_g = [x for x in gs if x not in ws][:6]
            only_w = [x for x in ws if x not in gs][:6]
            print(tname, nm, 'read', i, 'len', len(r), 'gpu-only', [(s, e, lst[k]) for s, e, k in only_g], 'oracle-only', [(s, e, lst[k]) for s, e, k in only_w])
            for s, e, k in (only_g + only_w)[:2]:
                print('   ctx', s, r[max(s - 30, 0):s].lower() + r[s:e + 30])
print('mismatching slices:', nbad)
