"""Batch boundary of the k-mer scan: drop-in for tg_tel.get_tel_repeat_comp_parallel (tg_tel.py:331-389) and
tg_tel.get_terminating_tl (tg_tel.py:156-235).

The reference runs gtrc_parallel_job (tg_tel.py:238-328) once per read inside a process pool; forked
workers cannot drive a GPU, so the pool is replaced by three phases over ALL reads:

  phase 1 (GPU)  pack, canonical base counts fwd/rev, orientation, density x2       tg_tel.py:262-266, 160-161
  phase 2 (GPU)  boundary calling: power signal, wavelet smoothing, region walk,
                 terminal-telomere scoring (tg_boundary.py); the filters of
                 tg_tel.py:278-305 as array arithmetic on the host                   tg_tel.py:156-235, 276-305
  phase 3 (GPU)  non-overlapping k-mer hits on every TVR+tel slice (and on whole
                 reads for interstitial-telomere candidates)                         tg_tel.py:287-290, 310

Same signature and the same 9-element return list (including the list-of-two-tuples shape of the
interstitial entries).  max_workers / max_pending are accepted and ignored.
With a callable `get_terminating_tl` (the reference's own function when the telomere-signal plots are
requested, or any custom boundary caller) phase 2 runs that function per read on the host instead; it
keeps calling get_telomere_kmer_density per read -- those calls are served from the densities
computed in phase 1 (tg_kmer.prime_density_cache).
"""
import numpy as np

import time

from . import tg_boundary, tg_kmer
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
    if get_terminating_tl is None and gtrc_params[13]:            # PLOT_TEL_SIGNALS: the plots are drawn by the reference's function
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
    lens = pr.lens.astype(np.int64)
    d2h = bc.nbytes
    oriented = [None] * n_reads                                     # the read as the reference sees it after :265-266

    def rdat_of(i):
        if oriented[i] is None:
            # (orient() above raised RC()'s KeyError for reads outside ACGTN)
            oriented[i] = rc_clean(all_read_dat[i][1]) if flip[i] else all_read_dat[i][1]
        return oriented[i]

    # ---- phase 2: boundary calling per read (tg_tel.py:276-279) ----
    t_ph.append(time.perf_counter())
    if get_terminating_tl is None:
        # on the device, from the window counts where they lie (tg_boundary.py)
        term = tg_boundary.terminating_tl_from_counts(d_p, d_q, pr.d_base_off, lens, TEL_WINDOW_SIZE, P_VS_Q_AMP_THRESH,
                                                      gtt_min_tel_score, 'q')
        d2h += term.nbytes
    else:
        # a caller-supplied get_terminating_tl (e.g. the reference's own, for the telomere-signal plots): per read on the
        # host; its get_telomere_kmer_density calls are served from the densities computed above
        h_p, h_q = pr.counts_to_host(d_p, d_q)
        cnt_p = pr.split_counts(h_p, TEL_WINDOW_SIZE)
        cnt_q = pr.split_counts(h_q, TEL_WINDOW_SIZE)
        d2h += 2 * int(getattr(pr, 'total', 0))
        term = np.full((n_reads, 4), -1, dtype=np.int64)
        for i in range(n_reads):
            my_rdat = rdat_of(i)
            if PLOT_TEL_SIGNALS:                                    # tg_tel.py:269-273 (pool passed the read index)
                tel_signal_plot_dat = (all_read_dat[i][0], TEL_SIGNAL_DIR + 'signal_' + str(i + index_offset).zfill(5) + '.png')
            else:
                tel_signal_plot_dat = None
            # densities as float64 arrays (every consumer only needs len(), indexing and numpy conversion: tg_kmer.py:106-115,
            # tg_plot.py:84-98); the all-zero second track of the reference (tg_kmer.py:67-76) is one shared read-only view
            n_out = len(cnt_p[i])
            zeros = np.broadcast_to(np.float64(0.0), (n_out,))
            tg_kmer.prime_density_cache(my_rdat, KMER_LIST, TEL_WINDOW_SIZE, (cnt_p[i].astype(np.float64) / TEL_WINDOW_SIZE, zeros))
            tg_kmer.prime_density_cache(my_rdat, KMER_LIST_REV, TEL_WINDOW_SIZE, (cnt_q[i].astype(np.float64) / TEL_WINDOW_SIZE, zeros))
            try:
                (tl, nontel, inter) = get_terminating_tl(my_rdat, 'q', gtt_params, telplot_dat=tel_signal_plot_dat)
            finally:
                tg_kmer.clear_density_cache()
            term[i, 0], term[i, 1] = tl, nontel
            if inter is not None:
                term[i, 2], term[i, 3] = inter[0], inter[1]
    termtel = np.minimum(term[:, 0].astype(np.int64), lens)        # tg_tel.py:278-279
    nontelend = np.minimum(term[:, 1].astype(np.int64), lens)
    ia, ib = term[:, 2].astype(np.int64), term[:, 3].astype(np.int64)
    # interstitial-telomere candidates (tg_tel.py:281-290): both anchors long enough
    want_i = (termtel == 0) & (ia >= 0) & (ib - ia >= MIN_INTERSTITIAL_TL) \
        & (np.minimum(ia, lens) >= MIN_FUSION_ANCHOR) & (np.maximum(lens - ib, 0) >= MIN_FUSION_ANCHOR)
    skipped_termtel = (termtel < FILT_TERM_TEL) if FILT_TERM_TEL > 0 else np.zeros(n_reads, dtype=bool)          # :292-298
    skipped_termunk = (nontelend > FILT_TERM_NONTEL) if FILT_TERM_NONTEL > 0 else np.zeros(n_reads, dtype=bool)
    skipped_termsub = (lens < termtel + FILT_TERM_SUBTEL) if FILT_TERM_SUBTEL > 0 else np.zeros(n_reads, dtype=bool)
    want_tvr = ~(skipped_termtel | skipped_termunk | skipped_termsub)
    tvr_lo = np.maximum(lens - termtel - FILT_TERM_SUBTEL, 0)       # tg_tel.py:301-305
    tvr_lo = np.where(lens - tvr_lo == 0, 0, tvr_lo)
    tvr_idx = np.nonzero(want_tvr)[0]
    i_idx = np.nonzero(want_i)[0]

    # ---- phase 3: GPU, all hit slices at once (reads are already oriented on the device) ----
    t_ph.append(time.perf_counter())
    hits_tvr = pr.hits(tvr_idx, tvr_lo[tvr_idx], lens[tvr_idx], KMER_LIST_REV, KMER_ISSUBSTRING)
    d2h += getattr(pr, 'last_hits_d2h', 0)
    hits_phases = dict(getattr(pr, 'last_hits_phases', {}))
    hits_ifwd = pr.hits(i_idx, np.zeros_like(i_idx), lens[i_idx], KMER_LIST, KMER_ISSUBSTRING)
    hits_irev = pr.hits(i_idx, np.zeros_like(i_idx), lens[i_idx], KMER_LIST_REV, KMER_ISSUBSTRING)
    t_ph.append(time.perf_counter())

    # ---- results in read order (tg_tel.py:352-375) ----
    kmer_hit_dat, interstitial_subtels = [], []
    interstitial_khd_by_readname, interstitial_khd_rev_by_readname = {}, {}
    meta = {'tvr': [], 'interstitial': []}
    tvr_len = (lens - tvr_lo).tolist()
    flip_l = flip.tolist()
    for q, i in enumerate(tvr_idx.tolist()):
        kmer_hit_dat.append([hits_tvr[q], tvr_len[i], 0, 'q', all_read_dat[i][0].split(' ')[0], DUMMY_TEL_MAPQ, rdat_of(i)])
        meta['tvr'].append((i + index_offset, flip_l[i]))
    for q, i in enumerate(i_idx.tolist()):
        my_rnm, my_rdat = all_read_dat[i][0], rdat_of(i)
        a, b = int(ia[i]), int(ib[i])
        interstitial_subtels.append([(f'interstitial-left_{len(my_rdat)}_{my_rnm}', my_rdat[:a]),       # list of two tuples, as
                                     (f'interstitial-right_{len(my_rdat)}_{my_rnm}', my_rdat[b:])])     # upstream
        meta['interstitial'].append((i + index_offset, flip_l[i]))
        interstitial_khd_by_readname[my_rnm] = [hits_ifwd[q], len(my_rdat), 0, 'q', my_rnm.split(' ')[0], DUMMY_TEL_MAPQ, my_rdat]
        interstitial_khd_rev_by_readname[my_rnm] = [hits_irev[q], len(my_rdat), 0, 'q', my_rnm.split(' ')[0], DUMMY_TEL_MAPQ, my_rdat]
    t_ph.append(time.perf_counter())
    LAST_STATS.clear()
    LAST_STATS.update(h2d_bytes=getattr(pr, 'h2d_bytes', 0), d2h_bytes=d2h,
                      phases={'gpu_pack_count_orient_density': t_ph[1] - t_ph[0], 'boundary_calling_and_filters': t_ph[2] - t_ph[1],
                              'gpu_hits_and_span_lists': t_ph[3] - t_ph[2], 'result_lists': t_ph[4] - t_ph[3]},
                      hits_phases=hits_phases,
                      boundary='device' if get_terminating_tl is None else 'callback')
    res = [kmer_hit_dat,
           termtel[tvr_idx].tolist(),
           nontelend[tvr_idx].tolist(),
           int(skipped_termtel.sum()),
           int(skipped_termunk.sum()),
           int(skipped_termsub.sum()),
           interstitial_subtels,
           interstitial_khd_by_readname,
           interstitial_khd_rev_by_readname]
    return (res, meta) if return_meta else res


def get_terminating_tl(rdat, pq, gtt_params, telplot_dat=None):
    """tg_tel.py:156-235 for one read (batch of one on the device) -> (tel_len, edge_nontel, largest interstitial or None)."""
    [SIGNAL_KMERS, SIGNAL_KMERS_REV, TEL_WINDOW_SIZE, P_VS_Q_AMP_THRESH, MIN_TEL_SCORE] = gtt_params
    pr = tg_kmer.PackedReads([rdat])
    d_p, d_q = pr.density_counts(SIGNAL_KMERS, SIGNAL_KMERS_REV, TEL_WINDOW_SIZE)
    n = max(len(rdat) - TEL_WINDOW_SIZE, 0)
    bb = tg_boundary.BoundaryBatch([n], TEL_WINDOW_SIZE)
    bb.power_from_counts(d_p, d_q, pr.d_base_off)
    bb.smooth()
    bb.regions(P_VS_Q_AMP_THRESH)
    (tl, edge, a, b) = (int(v) for v in bb.terminating(pq, MIN_TEL_SCORE)[0])
    largest = None if a < 0 else (a, b)
    if telplot_dat is not None:                                     # tg_tel.py:221-229; plotting is the reference's own code
        from source.tg_plot import plot_tel_signal
        (telplot_title, telplot_fn) = telplot_dat
        dp = (d_p[:n].cpu().numpy().astype(np.float64) / TEL_WINDOW_SIZE).tolist()
        dq = (d_q[:n].cpu().numpy().astype(np.float64) / TEL_WINDOW_SIZE).tolist()
        power = bb.signal_host(0) if n >= TEL_WINDOW_SIZE else []
        density_data = [dp, [0.0] * n, dq, [0.0] * n, power, len(rdat), TEL_WINDOW_SIZE]
        print(telplot_title, telplot_fn)
        plot_tel_signal(density_data, telplot_title, telplot_fn, tl_vals=[0, tl] if tl > 0 else None, interstitial_tels=[largest])
    return (tl, edge, largest)
