"""The reference's source/tg_align.py with the distance half bound to the B200 kernels (tg_align.py:109-189, 379-381).
get_scoring_matrix / get_aligner_object stay the reference's (the GPU path only reads the configured Bio.Align object), and so
does the MSA half (tg_align.py:192-376)."""
from source._overlay import exec_reference

exec_reference(__name__, globals())

from telogator2_b200.tg_align import get_dist_matrix_parallel, quick_compare_tvrs, tvr_distance  # noqa: E402,F401
