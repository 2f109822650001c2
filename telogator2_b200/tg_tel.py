"""Batch boundary of the k-mer scan: drop-in for tg_tel.get_tel_repeat_comp_parallel (tg_tel.py:331-389).

The reference runs gtrc_parallel_job (tg_tel.py:238-328) once per read inside a process pool; forked
workers cannot drive a GPU, so the pool is replaced by three phases over ALL reads:

  phase 1 (GPU)  pack, canonical base counts fwd/rev, orientation, density x2       tg_tel.py:262-266, 160-161
  phase 2 (CPU)  the reference's own get_terminating_tl / filters, per read          tg_tel.py:276-303 (unchanged code)
  phase 3 (GPU)  non-overlapping k-mer hits on every TVR+tel slice (and on whole
                 reads for interstitial-telomere candidates)                         tg_tel.py:287-290, 310

Same signature and the same 9-element return list (including the list-of-two-tuples shape of the
interstitial entries).  max_workers / max_pending are accepted and ignored.
get_terminating_tl is the reference's function (tg_tel.py:156-235, pywt wavelet code: outside the hot
path); it is taken from `source.tg_tel` when the reference package is importable, or passed in.  It
keeps calling get_telomere_kmer_density per read -- those calls are served from the densities
computed in phase 1 (tg_kmer.prime_density_cache).
"""
import numpy as np

import time

from . import tg_kmer
import gc

from .tg_util import rc_clean

LAST_STATS = {}          # byte counts and phase times of the most recent local batch (bench.py reads them)


def _reference_get_terminating_tl():
    try:
        from source.tg_tel import get_terminating_tl      # the reference package, when installed next to us
        return get_terminating_tl
    except Exception as e:                                 # noqa: BLE001
        raise ImportError('get_tel_repeat_comp_parallel needs the reference\'s get_terminating_tl '
                          '(source/tg_tel.py:156-235); pass it as get_terminating_tl=...') from e


def get_tel_repeat_comp_parallel(all_read_dat, gtrc_params, max_workers=4, max_pending=100, get_terminating_tl=None):
    """tg_tel.py:331-389.  Under an initialised torch.distributed process group (one process per GPU) the reads are split
    into contiguous shares of equal base count, every rank scans its share on its GPU, and the per-read results are
    exchanged (no read sequence travels: each rank re-attaches them from its own copy of all_read_dat), so every rank
    returns the complete 9-element list."""
    from .multi_gpu import _dist_state, shard_reads_by_bases, strip_tel_results, attach_tel_results
    dist, rank, world = _dist_state()
    if world <= 1 or len(all_read_dat) < world:
        return _tel_repeat_comp_local(all_read_dat, gtrc_params, get_terminating_tl, 0)
    bounds = shard_reads_by_bases([len(r[1]) for r in all_read_dat], world)
    lo, hi = bounds[rank], bounds[rank + 1]
    mine, meta = _tel_repeat_comp_local(all_read_dat[lo:hi], gtrc_params, get_terminating_tl, lo, return_meta=True)
    parts = [None] * world
    dist.all_gather_object(parts, strip_tel_results(mine, meta))
    return attach_tel_results(parts, all_read_dat)


def _tel_repeat_comp_local(all_read_dat, gtrc_params, get_terminating_tl=None, index_offset=0, return_meta=False):
    """The three phases on this process's GPU; index_offset = position of all_read_dat[0] in the caller's read list
    (the telomere-signal plot file names depend on it, tg_tel.py:269-273).  return_meta: also return, for every entry that
    carries a read sequence, (global read index, was it reverse-complemented) -- what the multi-GPU exchange ships instead."""
    [READ_TYPE, DUMMY_TEL_MAPQ,
     KMER_LIST, KMER_LIST_REV, KMER_ISSUBSTRING, CANONICAL_STRINGS, CANONICAL_STRINGS_REV,
     TEL_WINDOW_SIZE, P_VS_Q_AMP_THRESH,
     MIN_TEL_SCORE, FILT_TERM_TEL, FILT_TERM_NONTEL, FILT_TERM_SUBTEL,
     PLOT_TEL_SIGNALS, TEL_SIGNAL_DIR,
     MIN_INTERSTITIAL_TL, MIN_FUSION_ANCHOR] = gtrc_params
    if get_terminating_tl is None:
        get_terminating_tl = _reference_get_terminating_tl()
    # millions of small result lists are built below; the cycle collector only slows that down (4x on the list assembly)
    gc_was_on = gc.isenabled()
    gc.disable()
    try:
        return _tel_repeat_comp_phases(all_read_dat, gtrc_params, get_terminating_tl, index_offset, return_meta)
    finally:
        if gc_was_on:
            gc.enable()


def _tel_repeat_comp_phases(all_read_dat, gtrc_params, get_terminating_tl, index_offset, return_meta):
    [READ_TYPE, DUMMY_TEL_MAPQ,
     KMER_LIST, KMER_LIST_REV, KMER_ISSUBSTRING, CANONICAL_STRINGS, CANONICAL_STRINGS_REV,
     TEL_WINDOW_SIZE, P_VS_Q_AMP_THRESH,
     MIN_TEL_SCORE, FILT_TERM_TEL, FILT_TERM_NONTEL, FILT_TERM_SUBTEL,
     PLOT_TEL_SIGNALS, TEL_SIGNAL_DIR,
     MIN_INTERSTITIAL_TL, MIN_FUSION_ANCHOR] = gtrc_params
    gtt_min_tel_score = MIN_TEL_SCORE                               # tg_tel.py:248-251
    if FILT_TERM_TEL > 0:
        gtt_min_tel_score = min(FILT_TERM_TEL, MIN_TEL_SCORE)
    gtt_params = [KMER_LIST, KMER_LIST_REV, TEL_WINDOW_SIZE, P_VS_Q_AMP_THRESH, gtt_min_tel_score]
    n_reads = len(all_read_dat)

    # ---- phase 1: GPU ----
    t_ph = [time.perf_counter()]
    reads = [r[1] for r in all_read_dat]
    pr = tg_kmer.PackedReads(reads)
    bc = pr.basecount(CANONICAL_STRINGS, CANONICAL_STRINGS_REV).cpu().numpy()
    flip = bc[:, 0] > bc[:, 1]                                      # tg_tel.py:265
    pr.orient(flip)
    d_p, d_q = pr.density_counts(KMER_LIST, KMER_LIST_REV, TEL_WINDOW_SIZE)
    h_p, h_q = pr.counts_to_host(d_p, d_q)
    cnt_p = pr.split_counts(h_p, TEL_WINDOW_SIZE)
    cnt_q = pr.split_counts(h_q, TEL_WINDOW_SIZE)
    t_ph.append(time.perf_counter())

    # ---- phase 2: CPU, per read, the reference's own boundary calling and filters ----
    per_read = []
    slices = {'tvr': ([], [], []), 'ifwd': ([], [], []), 'irev': ([], [], [])}
    tel_signal_plot_num = 0
    for i in range(n_reads):
        (my_rnm, my_rdat, my_qdat) = all_read_dat[i]
        if flip[i]:
            my_rdat = rc_clean(my_rdat)                         # (orient() above raised RC()'s KeyError for reads outside ACGTN)
            if my_qdat is not None:
                my_qdat = my_qdat[::-1]
        if PLOT_TEL_SIGNALS:                                        # tg_tel.py:269-273 (pool passed the read index)
            tel_signal_plot_dat = (my_rnm, TEL_SIGNAL_DIR + 'signal_' + str(i + index_offset).zfill(5) + '.png')
            tel_signal_plot_num += 1
        else:
            tel_signal_plot_dat = None
        # densities as float64 arrays (every consumer only needs len(), indexing and numpy conversion: tg_kmer.py:106-115,
        # tg_plot.py:84-98); the all-zero second track of the reference (tg_kmer.py:67-76) is one shared read-only view
        n_out = len(cnt_p[i])
        zeros = np.broadcast_to(np.float64(0.0), (n_out,))
        tg_kmer.prime_density_cache(my_rdat, KMER_LIST, TEL_WINDOW_SIZE, (cnt_p[i].astype(np.float64) / TEL_WINDOW_SIZE, zeros))
        tg_kmer.prime_density_cache(my_rdat, KMER_LIST_REV, TEL_WINDOW_SIZE, (cnt_q[i].astype(np.float64) / TEL_WINDOW_SIZE, zeros))
        try:
            (my_terminating_tel, my_nontel_end, interstitial_tel) = get_terminating_tl(my_rdat, 'q', gtt_params,
                                                                                       telplot_dat=tel_signal_plot_dat)
        finally:
            tg_kmer.clear_density_cache()
        my_terminating_tel = min(my_terminating_tel, len(my_rdat))  # tg_tel.py:278-279
        my_nontel_end = min(my_nontel_end, len(my_rdat))
        st = {'rdat': my_rdat, 'rnm': my_rnm, 'isubtels': None, 'want_i': False, 'want_tvr': False,
              'termtel': my_terminating_tel, 'nontelend': my_nontel_end}
        if my_terminating_tel == 0 and interstitial_tel is not None and interstitial_tel[1] - interstitial_tel[0] >= MIN_INTERSTITIAL_TL:
            isubtel_left = my_rdat[:interstitial_tel[0]]            # tg_tel.py:281-290
            isubtel_right = my_rdat[interstitial_tel[1]:]
            if len(isubtel_left) >= MIN_FUSION_ANCHOR and len(isubtel_right) >= MIN_FUSION_ANCHOR:
                st['isubtels'] = [(f'interstitial-left_{len(my_rdat)}_{my_rnm}', isubtel_left),
                                  (f'interstitial-right_{len(my_rdat)}_{my_rnm}', isubtel_right)]
                st['want_i'] = True
                for key in ('ifwd', 'irev'):
                    slices[key][0].append(i); slices[key][1].append(0); slices[key][2].append(len(my_rdat))
        st['skipped_termtel'] = bool(FILT_TERM_TEL > 0 and my_terminating_tel < FILT_TERM_TEL)         # tg_tel.py:292-298
        st['skipped_termunk'] = bool(FILT_TERM_NONTEL > 0 and my_nontel_end > FILT_TERM_NONTEL)
        st['skipped_termsub'] = bool(FILT_TERM_SUBTEL > 0 and len(my_rdat) < my_terminating_tel + FILT_TERM_SUBTEL)
        if not (st['skipped_termtel'] or st['skipped_termunk'] or st['skipped_termsub']):
            my_subtel_end = max(len(my_rdat) - my_terminating_tel - FILT_TERM_SUBTEL, 0)                # tg_tel.py:301-305
            lo, hi = my_subtel_end, len(my_rdat)
            if hi - lo == 0:
                lo = 0
            st['want_tvr'] = True
            st['tvr_len'] = hi - lo
            slices['tvr'][0].append(i); slices['tvr'][1].append(lo); slices['tvr'][2].append(hi)
        per_read.append(st)

    # ---- phase 3: GPU, all hit slices at once (reads are already oriented on the device) ----
    t_ph.append(time.perf_counter())
    d2h = bc.nbytes + 2 * int(getattr(pr, 'total', 0))
    hits_tvr = pr.hits(*slices['tvr'], KMER_LIST_REV, KMER_ISSUBSTRING)
    d2h += getattr(pr, 'last_hits_d2h', 0)
    hits_ifwd = pr.hits(*slices['ifwd'], KMER_LIST, KMER_ISSUBSTRING)
    hits_irev = pr.hits(*slices['irev'], KMER_LIST_REV, KMER_ISSUBSTRING)
    t_ph.append(time.perf_counter())
    LAST_STATS.clear()
    LAST_STATS.update(h2d_bytes=getattr(pr, 'h2d_bytes', 0), d2h_bytes=d2h,
                      phases={'gpu_pack_count_orient_density_d2h': t_ph[1] - t_ph[0], 'host_per_read_boundary_calling': t_ph[2] - t_ph[1],
                              'gpu_hits_and_list_assembly': t_ph[3] - t_ph[2]})
    it_tvr, it_i = iter(hits_tvr), iter(zip(hits_ifwd, hits_irev))

    kmer_hit_dat, all_terminating_tl, all_nontel_end = [], [], []
    reads_removed_term_tel = reads_removed_term_unk = reads_removed_term_sub = 0
    interstitial_subtels = []
    interstitial_khd_by_readname, interstitial_khd_rev_by_readname = {}, {}
    meta = {'tvr': [], 'interstitial': []}
    for i, st in enumerate(per_read):                               # tg_tel.py:352-375, read order
        my_rdat, my_rnm = st['rdat'], st['rnm']
        if st['want_tvr']:
            kmer_hit_dat.append([next(it_tvr), st['tvr_len'], 0, 'q', my_rnm.split(' ')[0], DUMMY_TEL_MAPQ, my_rdat])
            all_terminating_tl.append(st['termtel'])
            all_nontel_end.append(st['nontelend'])
            meta['tvr'].append((i + index_offset, bool(flip[i])))
        if st['want_i']:
            (hf, hr) = next(it_i)
            interstitial_subtels.append(st['isubtels'])             # list of two tuples, as upstream
            meta['interstitial'].append((i + index_offset, bool(flip[i])))
            rn = all_read_dat[i][0]
            interstitial_khd_by_readname[rn] = [hf, len(my_rdat), 0, 'q', my_rnm.split(' ')[0], DUMMY_TEL_MAPQ, my_rdat]
            interstitial_khd_rev_by_readname[rn] = [hr, len(my_rdat), 0, 'q', my_rnm.split(' ')[0], DUMMY_TEL_MAPQ, my_rdat]
        reads_removed_term_tel += st['skipped_termtel'] * 1
        reads_removed_term_unk += st['skipped_termunk'] * 1
        reads_removed_term_sub += st['skipped_termsub'] * 1
    res = [kmer_hit_dat,
           all_terminating_tl,
           all_nontel_end,
           reads_removed_term_tel,
           reads_removed_term_unk,
           reads_removed_term_sub,
           interstitial_subtels,
           interstitial_khd_by_readname,
           interstitial_khd_rev_by_readname]
    return (res, meta) if return_meta else res
