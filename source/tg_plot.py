"""The reference's source/tg_plot.py, unchanged (outside the hot path)."""
from source._overlay import exec_reference

exec_reference(__name__, globals())
