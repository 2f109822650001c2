"""The two helpers of the reference's source/tg_util.py that sit on the hot path."""
import random

RC_DICT = {'A': 'T', 'C': 'G', 'G': 'C', 'T': 'A', 'N': 'N'}      # tg_util.py:41


_RC_TABLE = str.maketrans('ACGTN', 'TGCAN')
_RC_DELETE = {ord(c): None for c in 'ACGTN'}


def RC(s):
    """Reverse complement; KeyError on anything outside ACGTN (tg_util.py:433-434: ''.join(RC_DICT[n] for n in s[::-1])).
    Same result and the same KeyError, through str.translate (the per-letter dict loop runs 55 Mbases/s)."""
    out = s[::-1].translate(_RC_TABLE)
    if s.translate(_RC_DELETE):
        for n in s[::-1]:                          # the reference raises at the first offending letter of the reversed read
            RC_DICT[n]
    return out


def shuffle_seq(s):
    """tg_util.py:451-452 -- consumes the process-wide `random` stream."""
    return ''.join(random.sample(s, len(s)))
