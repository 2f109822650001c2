#!/usr/bin/env python
"""ALL bundled reads (83 ONT + 200 HiFi + 13 maize, BASELINE configs 1-3) through the REFERENCE's own scan functions
(build container only; needs /root/reference):

    python tests/golden/make_golden_all.py      ->  tests/golden/reads_all.json.gz, tests/golden/scan_all_golden.json.gz

reads_all.json.gz holds the reads themselves (name, sequence).  scan_all_golden.json.gz holds, per read, what
source/tg_kmer.py and the scan half of gtrc_parallel_job (tg_tel.py:262-266, 160-161, 310) compute:
  bc          get_telomere_base_count for CANONICAL_STRINGS and their reverse complements
  flip        the orientation decision
  dens_p/q    SHA-1 of the uint8 window counts (density * 100, rounded) of get_telomere_kmer_density on the ORIENTED read
              for KMER_LIST / KMER_LIST_REV, plus their sums
  hits        SHA-1 of the (k, start, end) triples of get_nonoverlapping_kmer_hits on the last 6000 bases of the oriented read
The reference modules are imported unmodified (absent third-party modules are inert stubs; none is touched here).
Digests keep the fixture small (the full arrays are ~12 MB); the GPU tests hash their own results the same way, and compare
them element-wise with the CPU oracle, which is held to the same digests in tests/test_oracle_golden.py.
"""
import gzip
import hashlib
import json
import os
import sys
from concurrent.futures import ProcessPoolExecutor
from unittest import mock

import numpy as np

REF = '/root/reference'
HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, REF)
for name in ['pywt', 'matplotlib', 'matplotlib.pyplot', 'matplotlib.colors', 'matplotlib.cm', 'matplotlib.lines',
             'matplotlib.patches', 'matplotlib.collections', 'Bio', 'Bio.Align', 'pysam']:
    sys.modules[name] = mock.MagicMock()

from source import tg_kmer, tg_util  # noqa: E402  (the reference)

SETS = [('human_ont', 'test_data/hg002-ont-1p.fa.gz', 'human_ont'),
        ('human_hifi', 'test_data/hg002-telreads_pacbio.sub.fa.gz', 'human_hifi'),
        ('maize_hifi', 'test_data/ZMMo17-hifi-7p8p.fa.gz', 'maize_hifi')]
W = 100


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


def sha(a):
    return hashlib.sha1(np.ascontiguousarray(a).tobytes()).hexdigest()


def hits_digest(hits):
    tri = np.array([(k, sp[0], sp[1]) for k, spans in enumerate(hits) for sp in spans], dtype=np.int32).reshape(-1, 3)
    return sha(tri), int(len(tri))


def one(job):
    (rdat, KL, ISSUB, CANON) = job
    KLR, CR = [tg_util.RC(k) for k in KL], [tg_util.RC(k) for k in CANON]
    f, r = tg_kmer.get_telomere_base_count(rdat, CANON), tg_kmer.get_telomere_base_count(rdat, CR)
    flip = f > r                                                   # tg_tel.py:265
    ori = tg_util.RC(rdat) if flip else rdat
    dp = np.rint(np.array(tg_kmer.get_telomere_kmer_density(ori, KL, W)[0]) * W).astype(np.uint8)
    dq = np.rint(np.array(tg_kmer.get_telomere_kmer_density(ori, KLR, W)[0]) * W).astype(np.uint8)
    hd, hn = hits_digest(tg_kmer.get_nonoverlapping_kmer_hits(ori[-6000:], KLR, ISSUB))
    return {'bc': [int(f), int(r)], 'flip': bool(flip), 'dens_p': sha(dp), 'dens_q': sha(dq),
            'sum_p': int(dp.astype(np.int64).sum()), 'sum_q': int(dq.astype(np.int64).sum()), 'hits': hd, 'n_spans': hn}


def main():
    tables = json.load(open(os.path.join(HERE, 'kmer_tables.json')))
    reads, golden = {}, {}
    for key, path, tkey in SETS:
        rs = read_fa(os.path.join(REF, path))
        T = tables[tkey]
        jobs = [(r, T['KMER_LIST'], T['KMER_ISSUBSTRING'], T['CANONICAL_STRINGS']) for _, r in rs]
        with ProcessPoolExecutor(max_workers=os.cpu_count()) as ex:
            res = list(ex.map(one, jobs, chunksize=2))
        reads[key] = [[n, r] for n, r in rs]
        golden[key] = res
        print(key, len(rs), 'reads', sum(len(r) for _, r in rs), 'bases', sum(x['flip'] for x in res), 'flipped')
    with gzip.open(os.path.join(HERE, 'reads_all.json.gz'), 'wt', compresslevel=9) as f:
        json.dump(reads, f)
    with gzip.open(os.path.join(HERE, 'scan_all_golden.json.gz'), 'wt', compresslevel=9) as f:
        json.dump(golden, f)


if __name__ == '__main__':
    main()
