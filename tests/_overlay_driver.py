"""Worker of tests/test_source_overlay.py (build container only: needs the reference checkout).

    python tests/_overlay_driver.py reference|overlay OUT.json

Runs the reference's OWN cluster_tvrs(clustering_only=True) and cluster_consensus_tvrs on the colour vectors / k-mer hits of
the bundled HiFi reads, either
  reference: over the reference's source/ package, with a scorer whose score() is the oracle's NW (Biopython is absent), or
  overlay:   over this repository's source/ overlay, i.e. through telogator2_b200.tg_align.get_dist_matrix_parallel -- the real
             host code (DistEngine, sharded_dist_matrix), with the two C entry points answered by the oracle through the real
             ctypes structs because the build container has no GPU (tests/test_dist_engine_host.py).
Absent third-party modules (pywt, matplotlib, Bio, pysam) are inert stubs in both arms.
"""
import gzip
import hashlib
import json
import os
import sys
import tempfile
from unittest import mock

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
REF = os.environ.get('TELOGATOR2_REFERENCE', '/root/reference')


def main():
    mode, out_path = sys.argv[1], sys.argv[2]
    for name in ['pywt', 'matplotlib', 'matplotlib.pyplot', 'matplotlib.pylab', 'matplotlib.colors', 'matplotlib.cm', 'matplotlib.lines',
                 'matplotlib.patches', 'matplotlib.collections', 'Bio', 'Bio.Align', 'pysam']:
        sys.modules[name] = mock.MagicMock()
    sys.path[:0] = [ROOT, HERE]
    import oracle
    if mode == 'reference':
        sys.path.insert(0, REF)
        from source import tg_align, tg_tvr
        assert os.path.dirname(os.path.abspath(tg_tvr.__file__)) == os.path.join(REF, 'source')
        get_sm = lambda canon, which_type='tvr': oracle.scoring_matrix(canon, which_type)        # noqa: E731
        get_al = lambda scoring_matrix=None, gap_bool=(True, True), which_type='tvr': oracle.ScoreStandInAligner(scoring_matrix, gap_bool)   # noqa: E731
    else:
        os.environ['TELOGATOR2_REFERENCE'] = REF
        from source import tg_align, tg_tvr
        assert os.path.dirname(os.path.abspath(tg_tvr.__file__)) == os.path.join(ROOT, 'source')
        from telogator2_b200 import _lib, tg_align as b200_align
        b200_align._BioAligner = b200_align._bio_sm = None              # (the stub above is not Biopython)
        assert tg_align.get_dist_matrix_parallel is b200_align.get_dist_matrix_parallel
        assert tg_tvr.get_dist_matrix_parallel is b200_align.get_dist_matrix_parallel       # the caller's own import line
        from test_dist_engine_host import OracleLib, _engine_from_scoring
        _lib.stream_ptr = lambda t: None
        b200_align.DistEngine = lambda sc: _engine_from_scoring(sc, OracleLib())
        get_sm, get_al = b200_align.get_scoring_matrix, b200_align.get_aligner_object       # Bio is a stub here
        # no GPU in the build container: the device linkage (bit-identical to scipy's: tests/test_gpu_linkage.py) is answered
        # by the oracle's restatement of scipy's algorithm, like the two C entry points of the distance matrix above
        from oracle import linkage as oracle_linkage
        assert tg_tvr.linkage.__module__ == 'source.tg_tvr' and tg_tvr._scipy_linkage.__module__.startswith('scipy')
        tg_tvr.linkage = lambda y, method='single', metric='euclidean', optimal_ordering=False: oracle_linkage.linkage(y, method)
    for m in (tg_align, tg_tvr):
        m.get_scoring_matrix, m.get_aligner_object = get_sm, get_al

    tables = json.load(open(os.path.join(HERE, 'golden', 'kmer_tables.json')))
    T = tables['human_hifi']
    meta = [T['KMER_LIST'], T['KMER_COLORS'], T['KMER_LETTER'], [tuple(f) for f in T['KMER_FLAGS']]]
    with gzip.open(os.path.join(HERE, 'golden', 'tvr_prep_golden.json.gz'), 'rt') as f:
        S = json.load(f)['sets']['human_hifi']
    K = len(meta[0])
    kmer_dat = []
    for e in S['kmer_dat'][:40]:
        hits = [[] for _ in range(K)]
        for k, a, b in e['hits']:
            hits[k].append([a, b])
        kmer_dat.append([hits, e['tlen'], 0, 'FWD', e['name'], 60, ''])
    out = {}
    with tempfile.TemporaryDirectory() as tmp:
        npz = os.path.join(tmp, 'dist.npz')
        res = tg_tvr.cluster_tvrs(kmer_dat, meta, 0.2, my_chr='chrUq', my_pos=0, aln_mode='ds', rand_shuffle_count=2, dist_in=npz,
                                  tvr_truncate=1000, num_processes=4, rng_seed=12345, clustering_only=True)
        out['cluster_tvrs'] = res[0]
        D = np.load(npz)['dist']
        out['npz_dtype'], out['npz_shape'] = str(D.dtype), list(D.shape)
        out['npz_sha1'] = hashlib.sha1(np.ascontiguousarray(D).tobytes()).hexdigest()
        # second call is served from the cache file the first one wrote (tg_tvr.py:155-161)
        res2 = tg_tvr.cluster_tvrs(kmer_dat, meta, 0.2, my_chr='chrUq', my_pos=0, aln_mode='ds', rand_shuffle_count=2, dist_in=npz,
                                   tvr_truncate=1000, num_processes=4, rng_seed=12345, clustering_only=True)
        out['cluster_tvrs_from_cache'] = res2[0]
        with gzip.open(os.path.join(HERE, 'golden', 'colorvecs.json.gz'), 'rt') as f:
            cv = json.load(f)['colorvecs']['human_hifi']
        cons = [c[:900 + 37 * i] for i, c in enumerate(cv[:14])]
        npz2 = os.path.join(tmp, 'cons.npz')
        out['cluster_consensus_tvrs'] = tg_tvr.cluster_consensus_tvrs(cons, meta, 0.05, dist_in=npz2, aln_mode='ms', gap_bool=(True, False),
                                                                      adjust_lens=False, rand_shuffle_count=3, linkage_method='single',
                                                                      num_processes=4, rng_seed=777)
        out['cons_sha1'] = hashlib.sha1(np.ascontiguousarray(np.load(npz2)['dist']).tobytes()).hexdigest()
    json.dump(out, open(out_path, 'w'))


if __name__ == '__main__':
    main()
