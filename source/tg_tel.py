"""The reference's source/tg_tel.py with get_tel_repeat_comp_parallel (tg_tel.py:331-389) bound to the three-phase GPU batch
and get_terminating_tl (tg_tel.py:156-235) to its device counterpart (telogator2_b200.tg_tel / tg_boundary).  With
PLOT_TEL_SIGNALS the batch calls the reference's own get_terminating_tl per read (it draws the plots)."""
from source._overlay import exec_reference

exec_reference(__name__, globals())

_reference_get_tel_repeat_comp_parallel = get_tel_repeat_comp_parallel  # noqa: F821  (kept for A/B runs)
_reference_get_terminating_tl = get_terminating_tl  # noqa: F821

from telogator2_b200.tg_tel import get_terminating_tl  # noqa: E402,F401


def get_tel_repeat_comp_parallel(all_read_dat, gtrc_params, max_workers=4, max_pending=100):
    from telogator2_b200 import tg_tel as _b200
    plots = bool(gtrc_params[13])                       # PLOT_TEL_SIGNALS
    return _b200.get_tel_repeat_comp_parallel(all_read_dat, gtrc_params, max_workers, max_pending,
                                              get_terminating_tl=_reference_get_terminating_tl if plots else None)
