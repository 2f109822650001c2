#!/usr/bin/env python
"""Pin the batch boundary of the scan against the REFERENCE's own driver (build container only; needs /root/reference):

    python tests/golden/make_golden_tel.py        ->  tests/golden/tel_batch_golden.json.gz

source/tg_tel.py is imported unmodified and its get_tel_repeat_comp_parallel (process pool + gtrc_parallel_job,
tg_tel.py:238-389) is run on the fixture reads.  get_terminating_tl (pywt wavelet code, outside the hot path) is replaced
by the same simple stand-in the GPU test uses (tests/test_gpu_tel_batch.make_gtt), driven by the reference's own
get_telomere_kmer_density.  Read sequences are stored as SHA-1 digests.
"""
import gzip
import hashlib
import json
import os
import sys
from unittest import mock

REF = '/root/reference'
HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, REF)
sys.path.insert(1, ROOT)
sys.path.insert(2, os.path.join(ROOT, 'tests'))
for name in ['pywt', 'matplotlib', 'matplotlib.pyplot', 'matplotlib.colors', 'matplotlib.cm', 'matplotlib.lines',
             'matplotlib.patches', 'matplotlib.collections', 'Bio', 'Bio.Align', 'pysam']:
    sys.modules[name] = mock.MagicMock()

from source import tg_kmer, tg_tel, tg_util   # noqa: E402  (the reference)
from test_gpu_tel_batch import make_gtt, tel_cases, digest_tel_result   # noqa: E402


def main():
    tables = json.load(open(os.path.join(HERE, 'kmer_tables.json')))
    with gzip.open(os.path.join(HERE, 'reads.json.gz'), 'rt') as f:
        reads_all = json.load(f)
    tg_tel.get_terminating_tl = make_gtt(tg_kmer.get_telomere_kmer_density)      # inherited by the forked pool workers
    out = {}
    for tkey, filters in tel_cases():
        T = tables[tkey]
        KL, ISSUB, CANON = T['KMER_LIST'], T['KMER_ISSUBSTRING'], T['CANONICAL_STRINGS']
        KLR, CANON_REV = [tg_util.RC(k) for k in KL], [tg_util.RC(k) for k in CANON]
        reads = [(name + ' extra', seq, None) for name, seq in reads_all[tkey]]
        reads += [(n, s, None) for n, s in reads_all['synthetic'] if len(s) > 0 and set(s) <= set('ACGTN')]
        params = [T['read_type'], 60, KL, KLR, ISSUB, CANON, CANON_REV, 100, 0.5, 100, filters[0], filters[1], filters[2],
                  False, '', 100, 1000]
        res = tg_tel.get_tel_repeat_comp_parallel(reads, params, max_workers=4)
        out[tkey] = digest_tel_result(res)
        print(tkey, len(res[0]), 'tvr entries,', len(res[6]), 'interstitial,', res[3:6], 'removed')
    with gzip.open(os.path.join(HERE, 'tel_batch_golden.json.gz'), 'wt') as f:
        json.dump(out, f)
    print('wrote tel_batch_golden.json.gz')


if __name__ == '__main__':
    main()
