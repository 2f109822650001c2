"""`source` overlay: the reference's package (zstephens/telogator2, source/) with its hot path bound to telogator2_b200.

Put this repository ahead of the reference checkout on sys.path (tools/run_telogator2.py does) and `telogator2.py`,
`make_telogator_ref.py` and every module of the reference run UNEDITED: each overlay module executes the reference's own
module source in its namespace (located through $TELOGATOR2_REFERENCE or a sibling checkout, see _overlay.py) and then
re-binds the hot-path functions to the B200 implementations:

  tg_kmer   get_telomere_base_count, get_telomere_kmer_density, get_nonoverlapping_kmer_hits
  tg_align  tvr_distance, get_dist_matrix_parallel, quick_compare_tvrs
  tg_tel    get_tel_repeat_comp_parallel (batch boundary of the scan; get_terminating_tl stays the reference's)
  tg_tvr    quick_get_tvrtel_lens
  tg_util   RC

Everything else -- boundary calling, MSA consensus, clustering, plotting, IO -- is the reference's code, which now reaches the
GPU through its own `from source.tg_align import ...` lines.
"""
