"""The reference's source/tg_tvr.py; cluster_tvrs / cluster_consensus_tvrs reach the GPU through the names they import from
source.tg_align.  quick_get_tvrtel_lens (tg_tvr.py:559-594) is bound to the batched colour-vector kernels."""
from source._overlay import exec_reference

exec_reference(__name__, globals())

from telogator2_b200.tg_tvr import quick_get_tvrtel_lens  # noqa: E402,F401
