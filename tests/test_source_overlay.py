"""The `source/` overlay: the reference's own callers -- cluster_tvrs(clustering_only=True) and
cluster_consensus_tvrs (tg_tvr.py:33-196, 336-446), unedited -- run over this repository's source/ package and must give the
cluster lists and the .npz cache bytes they give over the reference's own source/ package with the pinned scorer.
Build container only: it needs the reference checkout ($TELOGATOR2_REFERENCE or /root/reference); skipped elsewhere."""
import json
import os
import subprocess
import sys

import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.environ.get('TELOGATOR2_REFERENCE', '/root/reference')

pytestmark = pytest.mark.skipif(not os.path.isfile(os.path.join(REF, 'source', 'tg_tvr.py')), reason='reference checkout not present')


def _run(mode, tmp_path):
    out = tmp_path / f'{mode}.json'
    env = dict(os.environ, TELOGATOR2_REFERENCE=REF)
    subprocess.run([sys.executable, os.path.join(HERE, '_overlay_driver.py'), mode, str(out)], check=True, env=env, timeout=900)
    return json.load(open(out))


def test_reference_callers_over_the_overlay_equal_the_reference(tmp_path):
    ref = _run('reference', tmp_path)
    ovl = _run('overlay', tmp_path)
    assert ovl['npz_dtype'] == 'float32' and ovl['npz_shape'] == ref['npz_shape']
    assert ovl['npz_sha1'] == ref['npz_sha1'] and ovl['cons_sha1'] == ref['cons_sha1']          # .npz cache: byte compatible
    assert ovl['cluster_tvrs'] == ref['cluster_tvrs'] == ovl['cluster_tvrs_from_cache']
    assert ovl['cluster_consensus_tvrs'] == ref['cluster_consensus_tvrs']
    assert len(ref['cluster_tvrs']) > 3 and len(ref['cluster_consensus_tvrs']) > 3


def test_overlay_modules_bind_the_hot_path_and_keep_the_rest():
    """Every module the reference's telogator2.py imports resolves through the overlay; hot-path names are the B200 functions,
    the others are the reference's own code objects."""
    code = r'''
import os, sys
from unittest import mock
for name in ['pywt', 'matplotlib', 'matplotlib.pyplot', 'matplotlib.colors', 'matplotlib.cm', 'matplotlib.lines',
             'matplotlib.patches', 'matplotlib.collections', 'Bio', 'Bio.Align', 'pysam']:
    sys.modules[name] = mock.MagicMock()
sys.path.insert(0, %r)
import source
from source import tg_align, tg_kmer, tg_plot, tg_reader, tg_ref, tg_tel, tg_tvr, tg_util
import telogator2_b200.tg_align as A, telogator2_b200.tg_kmer as K, telogator2_b200.tg_tvr as V, telogator2_b200.tg_util as U
import telogator2_b200.tg_tel as T
ref = os.path.join(%r, 'source')
assert tg_align.get_dist_matrix_parallel is A.get_dist_matrix_parallel and tg_align.tvr_distance is A.tvr_distance
assert tg_align.quick_compare_tvrs is A.quick_compare_tvrs and tg_tvr.get_dist_matrix_parallel is A.get_dist_matrix_parallel
assert tg_kmer.get_telomere_kmer_density is K.get_telomere_kmer_density and tg_kmer.get_nonoverlapping_kmer_hits is K.get_nonoverlapping_kmer_hits
assert tg_kmer.get_telomere_base_count is K.get_telomere_base_count and tg_tvr.quick_get_tvrtel_lens is V.quick_get_tvrtel_lens
assert tg_util.RC is U.RC and tg_tel.get_tel_repeat_comp_parallel.__module__ == 'source.tg_tel'
# boundary calling (SURVEY 8f #3) is bound too; the reference's own functions stay reachable
assert tg_kmer.get_telomere_regions is K.get_telomere_regions and tg_kmer.wavelet_smooth is K.wavelet_smooth and tg_kmer.mad is K.mad
assert tg_tel.get_terminating_tl is T.get_terminating_tl
for fn in (tg_kmer._reference_get_telomere_regions, tg_kmer.read_kmer_tsv, tg_align.progressive_alignment, tg_align.get_final_tvr_consensus,
           tg_tel._reference_get_terminating_tl, tg_tel.get_allele_tsv_dat, tg_tvr.cluster_tvrs, tg_tvr.cluster_consensus_tvrs,
           tg_plot.plot_kmer_hits, tg_util.parse_read, tg_util.shuffle_seq):
    assert fn.__code__.co_filename.startswith(ref), fn
assert hasattr(tg_reader, 'TG_Reader') and hasattr(tg_ref, 'ReferenceFasta')
# the names telogator2.py imports (telogator2.py:14-20)
import re
src = open(os.path.join(%r, 'telogator2.py')).read()
for mod, names in re.findall(r'^from source\.(\w+)\s+import (.+)$', src, re.M):
    for nm in names.split(','):
        assert hasattr(getattr(source, mod), nm.strip()), (mod, nm)
print('ok')
''' % (os.path.dirname(HERE), REF, REF)
    r = subprocess.run([sys.executable, '-c', code], capture_output=True, text=True, timeout=300, env=dict(os.environ, TELOGATOR2_REFERENCE=REF))
    assert r.returncode == 0 and r.stdout.strip().endswith('ok'), r.stderr[-2000:]
