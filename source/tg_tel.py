"""The reference's source/tg_tel.py with get_tel_repeat_comp_parallel (tg_tel.py:331-389) bound to the three-phase GPU batch
(telogator2_b200.tg_tel).  gtrc_parallel_job's per-read logic -- get_terminating_tl and the filters -- stays the reference's."""
from source._overlay import exec_reference

exec_reference(__name__, globals())

_reference_get_tel_repeat_comp_parallel = get_tel_repeat_comp_parallel  # noqa: F821  (kept for A/B runs)


def get_tel_repeat_comp_parallel(all_read_dat, gtrc_params, max_workers=4, max_pending=100):
    from telogator2_b200 import tg_tel as _b200
    return _b200.get_tel_repeat_comp_parallel(all_read_dat, gtrc_params, max_workers, max_pending,
                                              get_terminating_tl=get_terminating_tl)  # noqa: F821  (the reference's, above)
