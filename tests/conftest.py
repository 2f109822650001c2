import gzip
import json
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, 'tests', 'golden')


def pytest_configure(config):
    config.addinivalue_line('markers', 'gpu: needs a CUDA device (run on the B200 box with -m gpu)')


@pytest.fixture(scope='session')
def kmer_tables():
    return json.load(open(os.path.join(GOLDEN, 'kmer_tables.json')))


@pytest.fixture(scope='session')
def golden_reads():
    with gzip.open(os.path.join(GOLDEN, 'reads.json.gz'), 'rt') as f:
        return json.load(f)


@pytest.fixture(scope='session')
def scan_golden():
    return dict(np.load(os.path.join(GOLDEN, 'scan_golden.npz')))


@pytest.fixture(scope='session')
def rng_golden():
    return dict(np.load(os.path.join(GOLDEN, 'rng_golden.npz')))


@pytest.fixture(scope='session')
def dp_golden():
    return dict(np.load(os.path.join(GOLDEN, 'dp_golden.npz')))


@pytest.fixture(scope='session')
def colorvecs():
    with gzip.open(os.path.join(GOLDEN, 'colorvecs.json.gz'), 'rt') as f:
        return json.load(f)


@pytest.fixture(scope='session')
def reads_all():
    """ALL bundled reads (83 ONT + 200 HiFi + 13 maize, BASELINE configs 1-3)."""
    with gzip.open(os.path.join(GOLDEN, 'reads_all.json.gz'), 'rt') as f:
        return json.load(f)


@pytest.fixture(scope='session')
def scan_all_golden():
    """Per read of reads_all: what the REFERENCE's scan functions compute (make_golden_all.py), as digests."""
    with gzip.open(os.path.join(GOLDEN, 'scan_all_golden.json.gz'), 'rt') as f:
        return json.load(f)


def sha1_of(a):
    import hashlib
    return hashlib.sha1(np.ascontiguousarray(a).tobytes()).hexdigest()


def hits_digest(hits):
    tri = np.array([(k, sp[0], sp[1]) for k, spans in enumerate(hits) for sp in spans], dtype=np.int32).reshape(-1, 3)
    return sha1_of(tri), int(len(tri))


RC_DICT = {'A': 'T', 'C': 'G', 'G': 'C', 'T': 'A', 'N': 'N'}


def rc_py(s):
    return ''.join(RC_DICT[c] for c in s[::-1])


def scan_cases(kmer_tables, golden_reads):
    """(tag, table, read) for every scan golden case, same enumeration as make_golden.py."""
    for rkey, rlist in golden_reads.items():
        tkeys = [rkey] if rkey != 'synthetic' else ['human_ont', 'human_hifi', 'maize_hifi']
        for tkey in tkeys:
            for ri, (rname, rdat) in enumerate(rlist):
                yield f'{rkey}|{tkey}|{ri}', kmer_tables[tkey], rdat


def hits_to_array(hits):
    k, s, e = [], [], []
    for ki, spans in enumerate(hits):
        for sp in spans:
            k.append(ki), s.append(sp[0]), e.append(sp[1])
    return np.stack([np.array(k, np.int32), np.array(s, np.int32), np.array(e, np.int32)]) if k else np.zeros((3, 0), np.int32)


@pytest.fixture(scope='session')
def boundary_golden():
    """Per read of reads_all: the REFERENCE's get_telomere_regions / get_terminating_tl outputs (make_golden_boundary.py)."""
    with gzip.open(os.path.join(GOLDEN, 'boundary_golden.json.gz'), 'rt') as f:
        return json.load(f)


def regions_digest(regs):
    cls = {None: 0, 'p': 1, 'q': 2}
    return sha1_of(np.array([(r[0], r[1], cls[r[2]]) for r in regs], dtype=np.int32).reshape(-1, 3))
