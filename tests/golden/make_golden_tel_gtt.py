#!/usr/bin/env python
"""The scan's batch boundary with the reference's OWN get_terminating_tl (build container only; needs /root/reference):

    python tests/golden/make_golden_tel_gtt.py        ->  tests/golden/tel_batch_gtt_golden.json.gz

source/tg_tel.get_tel_repeat_comp_parallel (process pool + gtrc_parallel_job, tg_tel.py:238-389) is run unmodified on ALL
bundled reads (tests/golden/reads_all.json.gz) with telogator2.py's default parameters.  PyWavelets is absent from the
image: `pywt` is oracle.boundary.fake_pywt() (see oracle/boundary.py for what that pins and what it does not).  Span lists
and read sequences are stored as SHA-1 digests (tests/test_gpu_boundary.compact_tel_result).
"""
import gzip
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
from oracle import boundary  # noqa: E402
for name in ['matplotlib', 'matplotlib.pyplot', 'matplotlib.colors', 'matplotlib.cm', 'matplotlib.lines',
             'matplotlib.patches', 'matplotlib.collections', 'Bio', 'Bio.Align', 'pysam']:
    sys.modules[name] = mock.MagicMock()
sys.modules['pywt'] = boundary.fake_pywt()

from source import tg_tel, tg_util   # noqa: E402  (the reference)
from test_gpu_boundary import compact_tel_result, gtt_cases, gtt_params_of   # noqa: E402


def main():
    tables = json.load(open(os.path.join(HERE, 'kmer_tables.json')))
    with gzip.open(os.path.join(HERE, 'reads_all.json.gz'), 'rt') as f:
        reads_all = json.load(f)
    out = {}
    for tkey, filters in gtt_cases():
        reads = [(name + ' extra', seq, None) for name, seq in reads_all[tkey]]
        params = gtt_params_of(tables[tkey], filters, tg_util.RC)
        res = tg_tel.get_tel_repeat_comp_parallel(reads, params, max_workers=16)
        out[tkey + '|' + ','.join(map(str, filters))] = compact_tel_result(res)
        print(tkey, filters, len(res[0]), 'tvr entries,', len(res[6]), 'interstitial,', res[3:6], 'removed')
    with gzip.open(os.path.join(HERE, 'tel_batch_gtt_golden.json.gz'), 'wt') as f:
        json.dump(out, f)
    print('wrote tel_batch_gtt_golden.json.gz', os.path.getsize(os.path.join(HERE, 'tel_batch_gtt_golden.json.gz')))


if __name__ == '__main__':
    main()
