"""The reference's source/tg_kmer.py with the scan functions bound to the B200 kernels (tg_kmer.py:66-102, 173-230).
read_kmer_tsv, get_canonical_letter, get_telomere_regions, mad and wavelet_smooth stay the reference's code."""
from source._overlay import exec_reference

exec_reference(__name__, globals())

from telogator2_b200.tg_kmer import (get_nonoverlapping_kmer_hits, get_telomere_base_count,  # noqa: E402,F401
                                     get_telomere_kmer_density)
