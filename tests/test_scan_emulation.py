"""CPU checks of the scan algorithms the CUDA kernels run (tests/scan_emul.py restates them position by position):
window-table coverage and the set-algebra form of get_nonoverlapping_kmer_hits against the oracle, and the lookup
tables the C library builds on the host against the Python construction."""
import ctypes as C
import gzip
import json
import os
import random

import numpy as np
import pytest

import oracle
import scan_emul as E

G = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden')
TABLES = json.load(open(os.path.join(G, 'kmer_tables.json')))
with gzip.open(os.path.join(G, 'reads.json.gz'), 'rt') as f:
    READS = json.load(f)


def rc(s):
    return s[::-1].translate(str.maketrans('ACGTN', 'TGCAN'))


def adversarial(rng, n):
    parts = ['CCCTAA', 'CCCCCC', 'CCCTCC', 'CCCCTACCCTAACCCCTA', 'CCCCTAACCCTAACCCCTAA', 'N', 'C', 'CC', 'TTAGGG', 'GGGGGG',
             'A', 'CCCTAAA', 'CCCTAAAAA', 'ACGT', 'GCCTAA']
    s = ''
    while len(s) < n:
        s += rng.choice(parts) * rng.randint(1, 4)
    return s[:n]


def cases(tname, rname):
    rng = random.Random(7)
    reads = [r[1][:1500] for r in READS[rname][:2]] + [r[1][-2500:] for r in READS[rname][:2]]
    reads += [adversarial(rng, 1200), rc(adversarial(rng, 700)), '', 'CCC', 'N' * 20 + 'CCCTAACCCTAA' + 'N' + 'CCCTAA', 'C' * 300]
    return reads


@pytest.mark.parametrize('tname,rname', [('human_hifi', 'human_hifi'), ('human_ont', 'human_ont'),
                                         ('maize_hifi', 'maize_hifi'), ('celegans_ont', 'synthetic')])
def test_window_table_algorithms_match_oracle(tname, rname):
    KL, ISS = TABLES[tname]['KMER_LIST'], TABLES[tname]['KMER_ISSUBSTRING']
    KR = [rc(k) for k in KL]
    tab = E.build_cov_table(KL, KR)
    hit_tabs = {id(KL): E.build_hit_table(KL), id(KR): E.build_hit_table(KR)}
    for rd in cases(tname, rname):
        for g, lst in enumerate((KL, KR)):
            cov = E.coverage_emulate(rd, lst, tab, g)
            assert sum(cov) == oracle.base_count(rd, lst)
            assert E.density_from_cov(cov, 100) == list(oracle.density_counts(rd, lst, 100))
            if rd:
                got = E.hits_emulate(rd, lst, ISS, *hit_tabs[id(lst)])
                assert got == oracle.nonoverlapping_hits(rd, lst, ISS)


@pytest.mark.parametrize('tname', ['human_hifi', 'human_ont', 'maize_hifi', 'celegans_ont'])
def test_library_pair_table_equals_python_construction(tname):
    from telogator2_b200 import _lib
    L = _lib.load()
    KL = TABLES[tname]['KMER_LIST']
    KR = [rc(k) for k in KL]
    blobs, offs = [], []
    for lst in (KL, KR):
        blobs.append(''.join(lst).encode())
        off = np.zeros(len(lst) + 1, dtype=np.int32)
        off[1:] = np.cumsum([len(k) for k in lst])
        offs.append(off)
    host = np.zeros(L.tg_kmer_table_pair_bytes(), dtype=np.uint8)
    _lib.check(L.tg_kmer_table_build_pair(blobs[0], offs[0].ctypes.data_as(C.c_void_p), len(KL),
                                          blobs[1], offs[1].ctypes.data_as(C.c_void_p), len(KR), host.ctypes.data_as(C.c_void_p)))
    tab = host[-4 * (1 << (2 * E.W)):].view(np.uint32)           # the table is the last member of the blob
    assert tab.tolist() == E.build_cov_table(KL, KR)
    short = E.build_cov_short_tables(KL, KR)
    n_short = 5464                                              # COV_SHORT_N
    got = host[-4 * (1 << (2 * E.W)) - 4 * n_short:-4 * (1 << (2 * E.W))].view(np.uint32)
    assert got[:len(short)].tolist() == short
    hdr = host[:32].view(np.int32)
    red = E.redundant(KL) + E.redundant(KR)
    n_bord = sum(E.is_bordered(k) and not r for k, r in zip(KL + KR, red))
    if tname.startswith('human'):
        assert sum(red) == 4                                     # the two HOR k-mers of both strands
        assert (hdr[6], hdr[7]) == (6, 6)                        # CCCCCC / CCCTCC and their reverse complements
    assert (hdr[1], hdr[2], hdr[3], hdr[4]) == (2 * len(KL), len(KL), n_bord, max(map(len, KL)))
