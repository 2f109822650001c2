"""telogator2_b200 -- B200-native hot path of Telogator2 (k-mer scan + TVR distance matrix).

Host side mirrors the reference's Python interface (source/tg_kmer.py, source/tg_align.py,
source/tg_tel.py:get_tel_repeat_comp_parallel); the work is done by hand-written sm_100a CUDA
kernels behind the C ABI in include/telogator2_b200.h.  There is no CPU fallback."""
__version__ = '0.1.0'
