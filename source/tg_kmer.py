"""The reference's source/tg_kmer.py with the scan functions (tg_kmer.py:66-102, 173-230) and the boundary calling
(get_telomere_regions, mad, wavelet_smooth: tg_kmer.py:105-170) bound to the B200 kernels.  read_kmer_tsv and
get_canonical_letter stay the reference's code; the reference's own boundary functions remain reachable as _reference_*."""
from source._overlay import exec_reference

exec_reference(__name__, globals())

_reference_get_telomere_regions = get_telomere_regions  # noqa: F821
_reference_wavelet_smooth = wavelet_smooth  # noqa: F821

from telogator2_b200.tg_kmer import (get_nonoverlapping_kmer_hits, get_telomere_base_count,  # noqa: E402,F401
                                     get_telomere_kmer_density, get_telomere_regions, mad, wavelet_smooth)
