"""Stage [0a] of the reference's main loop on the GPU: the telomere-read prefilter

    count_fwd = my_rdat.count(KMER_INITIAL)
    count_rev = my_rdat.count(KMER_INITIAL_RC)                      # telogator2.py:451-454
    keep = count_fwd >= MINIMUM_CANON_HITS or count_rev >= MINIMUM_CANON_HITS   # telogator2.py:459

for a batch of reads at once.  `str.count` semantics (non-overlapping occurrences, left to right, exact letters) are
reproduced bit for bit by csrc/tg_prefilter.cu.  No CPU fallback."""
import numpy as np

from . import _lib

ALIGN = 128


class AsciiReads:
    """Reads as ASCII on the device in the padded layout shared with tg_scan_pack (read starts at multiples of 128)."""

    def __init__(self, reads):
        torch = self.torch = _lib.require_cuda()
        self.L = _lib.load()
        self.n = len(reads)
        self.lens = np.array([len(r) for r in reads], dtype=np.int32)
        padded = (self.lens.astype(np.int64) + ALIGN - 1) // ALIGN * ALIGN
        self.base_off = np.zeros(self.n + 1, dtype=np.int64)
        np.cumsum(padded, out=self.base_off[1:])
        self.total = int(self.base_off[-1])
        tail = self.L.tg_prefilter_tail_bytes()
        host = torch.empty(self.total + tail, dtype=torch.uint8, pin_memory=True)
        buf = host.numpy()
        buf[:] = ord('N')
        for i, r in enumerate(reads):
            if self.lens[i]:
                o = int(self.base_off[i])
                buf[o:o + self.lens[i]] = np.frombuffer(r.encode('latin-1', 'replace'), dtype=np.uint8)
        self.d_ascii = host.cuda(non_blocking=True)
        self.d_base_off = torch.from_numpy(self.base_off).cuda()
        self.d_len = torch.from_numpy(self.lens).cuda()
        self.h2d_bytes = host.numel() + self.base_off.nbytes + self.lens.nbytes

    @classmethod
    def from_device(cls, d_ascii, base_off, lens):
        """Reads already resident on the device (d_ascii: uint8 tensor in the padded layout, tail included)."""
        self = cls.__new__(cls)
        self.torch = _lib.require_cuda()
        self.L = _lib.load()
        self.n = len(lens)
        self.lens = np.asarray(lens, dtype=np.int32)
        self.base_off = np.asarray(base_off, dtype=np.int64)
        self.total = int(self.base_off[-1])
        if d_ascii.numel() < self.total + self.L.tg_prefilter_tail_bytes():
            raise ValueError('device buffer lacks the tail padding (tg_prefilter_tail_bytes)')
        self.d_ascii = d_ascii
        self.d_base_off = self.torch.from_numpy(self.base_off).cuda()
        self.d_len = self.torch.from_numpy(self.lens).cuda()
        self.h2d_bytes = 0
        return self

    def count_pair_device(self, kmer_a, kmer_b):
        """int32 device tensors (count of kmer_a, count of kmer_b) per read."""
        torch = self.torch
        if len(kmer_a) != len(kmer_b) or not kmer_a:
            raise NotImplementedError('the device prefilter counts two non-empty patterns of the same length')
        if set(kmer_a + kmer_b) - set('ACGT'):
            raise NotImplementedError('the device prefilter needs upper-case ACGT patterns')
        ca = torch.zeros(max(self.n, 1), dtype=torch.int32, device='cuda')
        cb = torch.zeros(max(self.n, 1), dtype=torch.int32, device='cuda')
        if self.n:
            counter = torch.zeros(2, dtype=torch.int32, device='cuda')
            _lib.check(self.L.tg_prefilter_count(_lib.dptr(self.d_ascii), _lib.dptr(self.d_base_off), _lib.dptr(self.d_len), self.n,
                                                 kmer_a.encode('ascii'), kmer_b.encode('ascii'), len(kmer_a), _lib.dptr(ca), _lib.dptr(cb),
                                                 _lib.dptr(counter), _lib.stream_ptr(torch)))
        return ca[:self.n], cb[:self.n]


def count_initial_kmers(reads, kmer_initial, kmer_initial_rc):
    """(read.count(kmer_initial), read.count(kmer_initial_rc)) for every read, as two int64 arrays."""
    ar = AsciiReads(reads)
    ca, cb = ar.count_pair_device(kmer_initial, kmer_initial_rc)
    return ca.cpu().numpy().astype(np.int64), cb.cpu().numpy().astype(np.int64)


def tel_read_prefilter(reads, kmer_initial, kmer_initial_rc, minimum_canon_hits):
    """telogator2.py:451-459 for a batch: list of bool, True = the read goes to tel_reads.fa.gz."""
    cf, cr = count_initial_kmers(reads, kmer_initial, kmer_initial_rc)
    return [bool(a >= minimum_canon_hits or b >= minimum_canon_hits) for a, b in zip(cf, cr)]
