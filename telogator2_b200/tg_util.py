"""The two helpers of the reference's source/tg_util.py that sit on the hot path."""
import random

RC_DICT = {'A': 'T', 'C': 'G', 'G': 'C', 'T': 'A', 'N': 'N'}      # tg_util.py:41


def RC(s):
    """Reverse complement; KeyError on anything outside ACGTN (tg_util.py:433-434)."""
    return ''.join([RC_DICT[n] for n in s[::-1]])


def shuffle_seq(s):
    """tg_util.py:451-452 -- consumes the process-wide `random` stream."""
    return ''.join(random.sample(s, len(s)))
