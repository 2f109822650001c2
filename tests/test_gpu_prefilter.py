"""GPU parity of the stage [0a] prefilter (telogator2.py:451-459) against Python's own str.count -- which IS the
reference's code for this step."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _reads(golden_reads):
    rs = []
    for key in ('human_ont', 'human_hifi', 'maize_hifi', 'synthetic'):
        rs += [r[1] for r in golden_reads[key]]
    return rs


def _rc(s):
    return s[::-1].translate(str.maketrans('ACGT', 'TGCA'))


def test_counts_match_str_count(golden_reads, kmer_tables):
    from telogator2_b200 import tg_prefilter
    reads = _reads(golden_reads)
    rng = np.random.default_rng(4)
    # adversarial: overlapping repeats (CCCTAACCCTAA overlaps itself at shift 6), lower case, N, other letters that alias
    # the 2-bit code, matches straddling the 32-base lanes and the 992-base iterations, reads shorter than the pattern
    reads += ['CCCTAA' * 401, 'CCCTAA' * 3 + 'cccTAACCCTAA' + 'CCCTAACCCTAA', 'CCCTAACCCTAN' * 9, 'CCCTAAC', '', 'C',
              'ECCTAACCCTAA' * 5 + 'CCCTAACCCTAAA', 'A' * 985 + 'CCCTAACCCTAA' + 'G' * 7 + 'TTAGGGTTAGGG' * 3,
              'A' * 20 + 'CCCTAA' * 7 + 'A' * 961 + 'CCCTAA' * 5, 'TTAGGG' * 1000, 'AAAAAAAAAAAAAAAAAAAAAAAAAAAAAAA' * 40,
              ''.join(rng.choice(list('ACGT'), 40000))]
    for canon in ('CCCTAA', 'CCCTAAA', 'TTAGGC', 'AAAAAA', 'ACACAC', 'CCCTAACC'):
        ka, kb = canon + canon, _rc(canon) + _rc(canon)
        cf, cr = tg_prefilter.count_initial_kmers(reads, ka, kb)
        want_f = np.array([r.count(ka) for r in reads])
        want_r = np.array([r.count(kb) for r in reads])
        assert np.array_equal(cf, want_f), (canon, np.nonzero(cf != want_f)[0][:5])
        assert np.array_equal(cr, want_r), (canon, np.nonzero(cr != want_r)[0][:5])
    # patterns that are not a word twice (generic matching chain), odd lengths, the longest supported pattern
    for ka, kb in (('CCCTAACCCTAG', 'CTAGGGTTAGGG'), ('CCCTAAC', 'GTTAGGG'), ('C', 'G'), ('CCCTAA' * 5 + 'CC', 'GG' + 'TTAGGG' * 5),
                   ('TTAGGGTTAGGG', 'CCCTAACCCTAA')):
        cf, cr = tg_prefilter.count_initial_kmers(reads, ka, kb)
        assert np.array_equal(cf, np.array([r.count(ka) for r in reads])), ka
        assert np.array_equal(cr, np.array([r.count(kb) for r in reads])), kb
    keep = tg_prefilter.tel_read_prefilter(reads, 'CCCTAACCCTAA', 'TTAGGGTTAGGG', 5)
    assert keep == [(r.count('CCCTAACCCTAA') >= 5 or r.count('TTAGGGTTAGGG') >= 5) for r in reads]
    assert tg_prefilter.tel_read_prefilter([], 'CCCTAACCCTAA', 'TTAGGGTTAGGG', 5) == []
    with pytest.raises(NotImplementedError):
        tg_prefilter.count_initial_kmers(reads[:2], 'CCCTAN', 'NTAGGG')


def test_whole_genome_like_batch_properties(golden_reads):
    """A few hundred Mbases of mostly non-telomeric reads: counts of a sample against str.count, totals are stable."""
    from telogator2_b200 import tg_prefilter
    rng = np.random.default_rng(11)
    base = _reads(golden_reads)
    reads = [''.join(rng.choice(list('ACGT'), int(rng.integers(2000, 30000)))) for _ in range(300)] + base
    reads = reads * 4
    cf, cr = tg_prefilter.count_initial_kmers(reads, 'CCCTAACCCTAA', 'TTAGGGTTAGGG')
    for i in list(range(0, len(reads), 37)) + list(range(300, 300 + len(base))):
        assert cf[i] == reads[i].count('CCCTAACCCTAA') and cr[i] == reads[i].count('TTAGGGTTAGGG')
    n = len(reads) // 4
    assert np.array_equal(cf[:n], cf[n:2 * n]) and np.array_equal(cr[:n], cr[3 * n:])
