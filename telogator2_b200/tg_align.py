"""Drop-in for the distance half of the reference's source/tg_align.py (lines 14-189, 379-381).

Same names, signatures, argument meaning and error behaviour; the Needleman-Wunsch scores and the
MT19937 null model run on the GPU (see dist_engine.py).  The MSA half of the reference module
(tg_align.py:192-376: traceback alignments of <= 10 reads) is outside the hot path and stays the
reference's own code -- INTEGRATION.md shows the three-line binding.

When Biopython is importable, get_scoring_matrix / get_aligner_object return real Biopython objects
(the MSA half needs them); otherwise light stand-ins with the same attributes.  The GPU path only
reads attributes, so it accepts either.
"""
import random

import numpy as np

from .dist_engine import DistEngine, Scoring
from .tg_util import shuffle_seq

try:                                            # pragma: no cover - Biopython is absent in the build image
    from Bio.Align import PairwiseAligner as _BioAligner, substitution_matrices as _bio_sm
except Exception:                               # noqa: BLE001
    _BioAligner = _bio_sm = None

# tg_align.py:14-29
AMINO = ['A', 'C', 'D', 'E', 'F', 'G', 'H', 'I', 'K', 'L', 'M', 'N', 'P', 'Q', 'R', 'S', 'T', 'V', 'W', 'Y']
AMINO_2_IND = {AMINO[i]: i for i in range(len(AMINO))}
UNKNOWN = 'A'
GAP_CHR = '-'
MIN_VIABLE_SEQ_LEN = 1000
MAX_SEQ_DIST = 10.0
MIN_SEQ_DIST = 0.0001
IDEN_SCORE_SCALAR_CHEAT = 0.900
MAX_MSA_READCOUNT = 10


class SubstitutionArray(np.ndarray):
    """Stand-in for Bio.Align.substitution_matrices.Array (2-D, letter indexed) when Biopython is absent."""

    def __new__(cls, alphabet):
        obj = np.zeros((len(alphabet), len(alphabet)), dtype=np.float64).view(cls)
        obj.alphabet = alphabet
        return obj

    def __array_finalize__(self, obj):
        self.alphabet = getattr(obj, 'alphabet', None)

    def _ix(self, key):
        if isinstance(key, tuple):
            return tuple(self.alphabet.index(k) if isinstance(k, str) else k for k in key)
        return self.alphabet.index(key) if isinstance(key, str) else key

    def __getitem__(self, key):
        return np.ndarray.__getitem__(self, self._ix(key))

    def __setitem__(self, key, value):
        np.ndarray.__setitem__(self, self._ix(key), value)


class AlignerConfig:
    """Stand-in for a configured Bio.Align.PairwiseAligner (attributes only) when Biopython is absent."""

    def __init__(self):
        self.mode = 'global'
        self.match_score = 1.0
        self.mismatch_score = 0.0
        self.substitution_matrix = None
        self.open_gap_score = 0.0
        self.extend_gap_score = 0.0

    def __setattr__(self, name, value):
        object.__setattr__(self, name, value)
        if name in ('open_gap_score', 'extend_gap_score'):          # Biopython: sets every gap class at once
            kind = name.split('_')[0]
            for side in ('internal', 'left', 'right'):
                object.__setattr__(self, f'{kind}_{side}_gap_score', value)

    def score(self, seq_a, seq_b):
        eng = DistEngine(Scoring.from_aligner(self))
        return float(_single_scores(eng, [(seq_a, seq_b)])[0])


# value rules equivalent to replaying the assignments of tg_align.py:32-68 in order:
#   (canonical diagonal, canonical-vs-unknown, unknown-vs-gap, letter-vs-gap)
_MATRIX_TYPES = {'tvr': (0., -2., -2., -4.), 'consensus': (1., -2., -2., -4.),
                 'msa': (0., 1., -2., -4.), 'msa_refinement': (2., 1., 1., 0.)}


def get_scoring_matrix(canon_char, which_type='tvr'):
    """tg_align.py:32-68.  Note the unknown/unknown cell ends up -2 (the 2 is overwritten)."""
    if which_type not in _MATRIX_TYPES:
        print('Error: unknown which_type')
        exit(1)
    alphabet = ''.join(AMINO) + GAP_CHR
    m = _bio_sm.Array(alphabet, dims=2) if _bio_sm is not None else SubstitutionArray(alphabet)
    canon_diag, canon_unk, unk_gap, letter_gap = _MATRIX_TYPES[which_type]
    for c1 in AMINO:
        for c2 in AMINO:
            m[c1, c2] = 5. if c1 == c2 else -4.
        m[c1, UNKNOWN] = m[UNKNOWN, c1] = -2.
        m[c1, GAP_CHR] = m[GAP_CHR, c1] = letter_gap
    m[GAP_CHR, GAP_CHR] = 0.
    m[canon_char, canon_char] = canon_diag
    m[canon_char, UNKNOWN] = m[UNKNOWN, canon_char] = canon_unk
    m[UNKNOWN, GAP_CHR] = m[GAP_CHR, UNKNOWN] = unk_gap
    return m


def get_aligner_object(scoring_matrix=None, gap_bool=(True, True), which_type='tvr'):
    """tg_align.py:71-106."""
    if which_type == 'tvr':
        gap_open, gap_extend = -4., -4.
    elif which_type == 'msa':
        gap_open, gap_extend = -16., -4.
    else:
        print('Error: unknown which_type')
        exit(1)
    aligner = _BioAligner() if _BioAligner is not None else AlignerConfig()
    aligner.mode = 'global'
    aligner.match_score = 5.
    aligner.mismatch_score = -4.
    aligner.open_gap_score = gap_open
    aligner.extend_gap_score = gap_extend
    if scoring_matrix is not None:
        aligner.substitution_matrix = scoring_matrix
    for side, apply_penalty in zip(['left', 'right'], gap_bool):
        o, e = (gap_open, gap_extend) if apply_penalty else (0.0, 0.0)
        try:                                                # Biopython >= 1.86 names, then 1.82-1.85 names
            setattr(aligner, f'open_{side}_gap_score', o)
            setattr(aligner, f'extend_{side}_gap_score', e)
        except AttributeError:
            setattr(aligner, f'{side}_open_gap_score', o)
            setattr(aligner, f'{side}_extend_gap_score', e)
    return aligner


def _single_scores(engine, seq_pairs):
    """NW scores of explicit (a, b) string pairs on the GPU."""
    from .dist_engine import pack_jobs
    seqs = [s for ab in seq_pairs for s in ab]
    codes, off, sub = engine.sc.encode_all(seqs)
    n = np.diff(off).astype(np.int32)
    k = len(seq_pairs)
    jobs = pack_jobs(off[0:2 * k:2], off[1:2 * k:2], n[0::2], n[1::2], np.arange(k, dtype=np.int32), True)
    return engine.run_jobs(engine._to_dev(codes), jobs, sub, k)


def tvr_distance(tvr_i, tvr_j, aligner, adjust_lens=True, min_viable=True, randshuffle=3, rng_seed=None):
    """tg_align.py:109-151, one pair.  Like the reference it (re)seeds and consumes the process-wide
    `random` stream, so stage [7] (quick_compare_tvrs, no reseed) sees the same stream as upstream;
    the alignments themselves run on the GPU."""
    if rng_seed is not None:
        random.seed(rng_seed)
        np.random.seed(rng_seed)
        np.random.default_rng(rng_seed)
    if adjust_lens:
        min_len = min(len(tvr_i), len(tvr_j))
        if min_viable:
            min_len = max(min_len, MIN_VIABLE_SEQ_LEN)
        seq_i, seq_j = tvr_i[:min_len], tvr_j[:min_len]
    else:
        seq_i, seq_j = tvr_i, tvr_j
    sc = Scoring.from_aligner(aligner)
    eng = DistEngine(sc)
    aln_score = float(_single_scores(eng, [(seq_i, seq_j)])[0])
    if sc.alphabet is None:
        iden_score = sc.match * ((len(seq_i) + len(seq_j)) / 2.)
    else:
        diag = {ch: float(sc.sub[i, i]) for i, ch in enumerate(sc.alphabet)}
        iden_score = (sum(diag[c] for c in seq_i) + sum(diag[c] for c in seq_j)) / 2.
    if aln_score < -(IDEN_SCORE_SCALAR_CHEAT * iden_score):
        return MAX_SEQ_DIST
    shuffled = [(shuffle_seq(seq_i), shuffle_seq(seq_j)) for _ in range(randshuffle)]
    rand_scores = [float(v) for v in _single_scores(eng, shuffled)] if shuffled else []
    rand_score_shuffle = np.mean(rand_scores)
    if rand_score_shuffle >= aln_score:
        dist_out = MAX_SEQ_DIST
    else:
        dist_out = min(-np.log((aln_score - rand_score_shuffle) / (iden_score - rand_score_shuffle)), MAX_SEQ_DIST)
    dist_out = max(dist_out, MIN_SEQ_DIST)
    return dist_out


def get_dist_matrix_parallel(sequences, aligner, adjust_lens, min_viable, rand_shuffle_count, max_workers=4,
                             max_pending=100, tasks_per_worker=5000, rng_seed=12345, print_progress=False):
    """tg_align.py:154-189.  max_workers / max_pending / tasks_per_worker are accepted and ignored
    (process-pool knobs); the result never depended on them.  float32 (n, n), symmetric, zero diagonal,
    raw distances in [1e-4, 10] (the caller divides by MAX_SEQ_DIST, tg_tvr.py:154)."""
    eng = DistEngine(Scoring.from_aligner(aligner))
    from .multi_gpu import sharded_dist_matrix
    return sharded_dist_matrix(eng, sequences, adjust_lens, min_viable, rand_shuffle_count, rng_seed)


def get_dist_matrices_parallel(sequence_lists, aligner, adjust_lens, min_viable, rand_shuffle_count, rng_seed=12345):
    """[get_dist_matrix_parallel(seqs, aligner, adjust_lens, min_viable, rand_shuffle_count, rng_seed=rng_seed) for seqs in
    sequence_lists] in one device pass (no upstream counterpart: upstream calls tg_align.py:154-189 once per cluster,
    tg_tvr.py:147-159 / 375-390).  Small matrices leave most of a GPU idle one at a time; batched they share it."""
    eng = DistEngine(Scoring.from_aligner(aligner))
    from .multi_gpu import sharded_dist_matrices
    return sharded_dist_matrices(eng, sequence_lists, adjust_lens, min_viable, rand_shuffle_count, rng_seed)


def quick_compare_tvrs(tvr_1, tvr_2):
    """tg_align.py:379-381 (match/mismatch scoring, free right end gaps, no reseed)."""
    aligner = get_aligner_object(gap_bool=(True, False), which_type='tvr')
    return tvr_distance(tvr_1, tvr_2, aligner, adjust_lens=False, min_viable=False) / MAX_SEQ_DIST
