"""The two helpers of the reference's source/tg_util.py that sit on the hot path."""
import random

RC_DICT = {'A': 'T', 'C': 'G', 'G': 'C', 'T': 'A', 'N': 'N'}      # tg_util.py:41


_RC_BYTES = bytes.maketrans(b'ACGTN', b'TGCAN')


def rc_clean(s):
    """Reverse complement of a read already known to hold only ACGTN (the pack kernel flags the others)."""
    return s.encode().translate(_RC_BYTES)[::-1].decode()


def RC(s):
    """Reverse complement; KeyError on anything outside ACGTN (tg_util.py:433-434: ''.join(RC_DICT[n] for n in s[::-1])).
    Same result and the same KeyError, through bytes.translate (the per-letter dict loop runs 55 Mbases/s)."""
    b = s.encode()
    if b.translate(None, b'ACGTN'):
        for n in s[::-1]:                          # the reference raises at the first offending letter of the reversed read
            RC_DICT[n]
    return b.translate(_RC_BYTES)[::-1].decode()


def shuffle_seq(s):
    """tg_util.py:451-452 -- consumes the process-wide `random` stream."""
    return ''.join(random.sample(s, len(s)))
