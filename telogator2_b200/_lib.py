"""ctypes binding of libtg_b200.so (include/telogator2_b200.h).  Fails loudly: no library or no
CUDA device means the hot path is unavailable -- there is no CPU fallback."""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
SO_PATH = os.path.join(_HERE, 'libtg_b200.so')

TG_OK, TG_EINVAL, TG_ECUDA, TG_ERANGE, TG_EOVERFLOW = 0, -1, -2, -3, -4
MAXA = 32

# struct tg_nw_job (64 bytes) / tg_shuffle_job (40 bytes) as numpy record dtypes
NW_JOB = np.dtype([('row_off', '<i8', (2,)), ('col_off', '<i8', (2,)), ('n', '<i4', (2,)), ('m', '<i4', (2,)),
                   ('out', '<i4', (2,)), ('pad_', '<i4', (2,))], align=True)
SHUFFLE_JOB = np.dtype([('seed', '<u8'), ('src_i', '<i8'), ('src_j', '<i8'), ('dst_off', '<i8'),
                        ('n', '<i4'), ('m', '<i4')], align=True)
assert NW_JOB.itemsize == 64 and SHUFFLE_JOB.itemsize == 40


class NwScoring(C.Structure):
    _fields_ = [('alphabet', C.c_int32), ('gap', C.c_int32), ('gap_left', C.c_int32), ('gap_right', C.c_int32),
                ('sub', C.c_int16 * (MAXA * MAXA))]


class DistBatchArgs(C.Structure):
    _fields_ = [('d_pool', C.c_void_p), ('d_seq_off', C.c_void_p), ('d_order', C.c_void_p), ('d_diag_cs', C.c_void_p),
                ('d_aln', C.c_void_p), ('d_rnd', C.c_void_p), ('d_shape', C.c_void_p), ('d_flags', C.c_void_p),
                ('d_iden', C.c_void_p), ('d_scratch', C.c_void_p), ('d_counter', C.c_void_p), ('d_mt_work', C.c_void_p),
                ('q0', C.c_int64), ('q1', C.c_int64), ('seed0', C.c_int64), ('shuf_base', C.c_int64), ('slot', C.c_int64),
                ('match', C.c_double),
                ('n_seq', C.c_int32), ('adjust_lens', C.c_int32), ('min_viable', C.c_int32), ('R', C.c_int32),
                ('max_len', C.c_int32), ('n_mat', C.c_int32), ('d_mat_pair_off', C.c_void_p), ('d_mat_seq_off', C.c_void_p),
                ('d_seq_mask', C.c_void_p), ('prof_rows', C.c_int32), ('phases', C.c_int32), ('smem_reserve', C.c_int32),
                ('pad_', C.c_int32)]


class DistEpilogueArgs(C.Structure):
    _fields_ = [('d_seq_off', C.c_void_p), ('d_order', C.c_void_p), ('d_aln', C.c_void_p), ('d_rnd', C.c_void_p),
                ('d_flags', C.c_void_p), ('d_iden', C.c_void_p), ('d_matrix', C.c_void_p), ('d_ambig', C.c_void_p),
                ('d_stats', C.c_void_p), ('q0', C.c_int64), ('q1', C.c_int64), ('ambig_base', C.c_int64),
                ('n_seq', C.c_int32), ('adjust_lens', C.c_int32), ('min_viable', C.c_int32), ('R', C.c_int32),
                ('ambig_cap', C.c_int32), ('margin_log2', C.c_int32), ('n_mat', C.c_int32), ('pad_', C.c_int32),
                ('d_mat_pair_off', C.c_void_p), ('d_mat_seq_off', C.c_void_p), ('d_mat_out_off', C.c_void_p)]


class TgError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f'telogator2_b200 error {code}: {msg}')
        self.code = code


_lib = None

EXPORTS = ['tg_last_error', 'tg_version', 'tg_device_sm_count',
           'tg_nw_wave_scratch_bytes', 'tg_nw_wave_check', 'tg_nw_scores_wave', 'tg_nw_scores_wave32', 'tg_nw_generic_max_m',
           'tg_nw_scores_generic', 'tg_mt_shuffle', 'tg_dist_batch', 'tg_dist_batch_mt_work_bytes', 'tg_dist_epilogue', 'tg_wave_timing', 'tg_wave_timing_read', 'tg_int_alu_peak',
           'tg_kmer_table_bytes', 'tg_kmer_table_build', 'tg_kmer_table_pair_bytes', 'tg_kmer_table_build_pair', 'tg_scan_pack', 'tg_scan_basecount', 'tg_scan_orient',
           'tg_scan_density', 'tg_scan_hits',
           'tg_cv_paint', 'tg_cv_density_boundary', 'tg_cv_lead_run', 'tg_cv_gather', 'tg_cv_denoise_tile', 'tg_cv_denoise',
           'tg_prefilter_tail_bytes', 'tg_prefilter_count',
           'tg_bound_power', 'tg_bound_smooth', 'tg_bound_regions', 'tg_bound_terminating',
           'tg_linkage_expand', 'tg_linkage_widen', 'tg_linkage_run', 'tg_linkage_multi_work_bytes', 'tg_linkage_run_multi']


def load():
    """Load the CUDA library (no device needed to load; compute calls need one)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(SO_PATH):
        raise ImportError(f'{SO_PATH} is missing: build it with `python -m telogator2_b200.build` '
                          '(nvcc, sm_100a). telogator2_b200 has no CPU fallback.')
    L = C.CDLL(SO_PATH)
    vp, i32, i64, sz = C.c_void_p, C.c_int, C.c_int64, C.c_size_t
    L.tg_last_error.restype = C.c_char_p
    L.tg_nw_wave_scratch_bytes.restype = sz
    L.tg_nw_wave_scratch_bytes.argtypes = [i32]
    L.tg_nw_wave_check.argtypes = [C.POINTER(NwScoring), vp, i64]
    L.tg_nw_scores_wave.argtypes = [vp, vp, i64, i32, C.POINTER(NwScoring), vp, vp, vp, vp]
    L.tg_nw_scores_wave32.argtypes = [vp, vp, i64, i32, C.POINTER(NwScoring), vp, vp, vp, vp]
    L.tg_nw_scores_generic.argtypes = [vp, vp, i64, i32, C.POINTER(NwScoring), vp, vp]
    L.tg_mt_shuffle.argtypes = [vp, vp, i64, i32, i32, vp]
    L.tg_dist_batch.argtypes = [C.POINTER(DistBatchArgs), C.POINTER(NwScoring), vp]
    L.tg_dist_epilogue.argtypes = [C.POINTER(DistEpilogueArgs), vp]
    L.tg_wave_timing.argtypes = [i32]
    L.tg_wave_timing_read.argtypes = [C.POINTER(C.c_double), C.POINTER(C.c_int)]
    L.tg_dist_batch_mt_work_bytes.restype = sz
    L.tg_dist_batch_mt_work_bytes.argtypes = [i64, i32, i32]
    L.tg_int_alu_peak.argtypes = [i32, C.POINTER(C.c_double), C.POINTER(C.c_double)]
    L.tg_device_sm_count.argtypes = [C.POINTER(C.c_int)]
    L.tg_kmer_table_bytes.restype = sz
    L.tg_kmer_table_build.argtypes = [C.c_char_p, vp, i32, vp, vp, vp]
    L.tg_scan_pack.argtypes = [vp, vp, vp, i32, vp, vp, vp, vp]
    L.tg_kmer_table_build_pair.argtypes = [C.c_char_p, vp, i32, C.c_char_p, vp, i32, vp]
    L.tg_scan_basecount.argtypes = [vp, vp, i64, vp, i32, vp, vp, vp]
    L.tg_kmer_table_pair_bytes.restype = sz
    L.tg_scan_orient.argtypes = [vp, vp, vp, vp, vp, i32, vp, vp, vp]
    L.tg_scan_density.argtypes = [vp, vp, i64, vp, i32, vp, vp, vp]
    L.tg_scan_hits.argtypes = [vp, vp, vp, vp, vp, vp, i32, vp, vp, vp, vp, vp, i64, vp, vp]
    L.tg_cv_paint.argtypes = [vp, vp, vp, vp, i32, vp, i32, i32, vp, vp]
    L.tg_cv_density_boundary.argtypes = [vp, vp, vp, vp, i32, C.c_char_p, i32, i32, C.c_double, i32, i32, vp, vp, vp]
    L.tg_cv_lead_run.argtypes = [vp, vp, vp, vp, vp, i32, i32, vp, vp]
    L.tg_cv_gather.argtypes = [vp, vp, vp, vp, i64, vp, vp]
    L.tg_cv_denoise.argtypes = [vp, vp, vp, vp, vp, vp, i32, i64, i32, i32, i32, C.c_char_p, i32, C.c_char_p, i32, vp, vp]
    L.tg_prefilter_count.argtypes = [vp, vp, vp, i32, C.c_char_p, C.c_char_p, i32, vp, vp, vp, vp]
    f64 = C.c_double
    L.tg_bound_power.argtypes = [vp, vp, vp, vp, vp, i32, i32, i64, i32, vp, vp]
    L.tg_bound_smooth.argtypes = [vp, vp, vp, vp, vp, vp, vp, i32, i32, i64, i32, i32, f64, vp, vp, vp, vp, vp, vp, vp, vp]
    L.tg_bound_regions.argtypes = [vp, vp, vp, vp, i32, i32, f64, vp, vp, vp, vp, vp, vp, vp]
    L.tg_bound_terminating.argtypes = [vp, vp, vp, vp, vp, i32, i32, i32, f64, vp, vp]
    L.tg_linkage_expand.argtypes = [vp, i32, vp, vp]
    L.tg_linkage_widen.argtypes = [vp, i32, C.c_float, vp, vp]
    L.tg_linkage_run.argtypes = [vp, i32, C.c_char_p, vp, vp, vp, vp, vp]
    L.tg_linkage_multi_work_bytes.restype = sz
    L.tg_linkage_multi_work_bytes.argtypes = [i32, i32]
    L.tg_linkage_run_multi.argtypes = [vp, i32, C.c_char_p, i32, vp, vp, vp, vp, vp]
    _lib = L
    return L


def check(rc):
    if rc != TG_OK:
        raise TgError(rc, load().tg_last_error().decode('utf-8', 'replace'))


def require_cuda():
    """torch + a CUDA device, or a loud failure."""
    import torch
    if not torch.cuda.is_available():
        raise RuntimeError('telogator2_b200 needs a CUDA device (B200, sm_100a); there is no CPU fallback')
    load()
    return torch


def stream_ptr(torch):
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def dptr(t):
    return C.c_void_p(t.data_ptr()) if t is not None else C.c_void_p(0)
