"""Worker of tests/test_multi_gpu_gloo.py: one rank of a world_size-2 gloo job exercising the sharding +
gather logic of telogator2_b200.multi_gpu with the CPU oracle standing in for the device call."""
import gzip
import json
import os
import sys

import numpy as np
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import oracle  # noqa: E402
from telogator2_b200.dist_engine import DistEngine, NOT_COMPUTED  # noqa: E402
from telogator2_b200.dist_engine import fast_ranges  # noqa: E402
from telogator2_b200.multi_gpu import shard_pair_indices, sharded_dist_matrix, sharded_pair_scores  # noqa: E402


def main():
    out_path = sys.argv[1]
    dist.init_process_group('gloo')
    rank, world = dist.get_rank(), dist.get_world_size()
    with gzip.open(os.path.join(ROOT, 'tests', 'golden', 'colorvecs.json.gz'), 'rt') as f:
        cv = json.load(f)['colorvecs']['human_ont']
    seqs = [c[:400 + 13 * i] for i, c in enumerate(cv[:14])]
    P = oracle.AlignerParams(oracle.scoring_matrix('C', 'tvr'), (True, True))
    R, seed = 2, 4321
    n = len(seqs)
    n_pairs = n * (n - 1) // 2
    _D, aln_full, rnd_full, _cells = oracle.dist_matrix_c(seqs, P, True, True, R, seed, want_scores=True)

    def compute(shard):
        idx = shard_pair_indices(n_pairs, shard[0], shard[1], block=8)
        aln = np.full(n_pairs, NOT_COMPUTED, dtype=np.int32)
        rnd = np.full((n_pairs, R), NOT_COMPUTED, dtype=np.int32)
        iden = np.zeros(n_pairs)
        nn = np.zeros(n_pairs, dtype=np.int32)
        mm = np.zeros(n_pairs, dtype=np.int32)
        pairs = [(i, j) for i in range(n) for j in range(i + 1, n)]
        for p in idx:
            i, j = pairs[p]
            det = {}
            oracle.tvr_distance_py(seqs[i], seqs[j], P, True, True, R, seed + int(p), detail=det)
            aln[p] = int(det['aln'])
            iden[p] = det['iden']
            nn[p], mm[p] = det['n'], det['m']
            for k, v in enumerate(det['rand']):
                rnd[p, k] = int(v)
        return dict(aln=aln, rnd=rnd, iden=iden, n=nn, m=mm, order=None)

    res = sharded_pair_scores(None, seqs, True, True, R, seed, compute=compute, device='cpu')
    assert np.array_equal(res['aln'], aln_full), 'gathered aln differs'
    assert np.array_equal(res['rnd'], rnd_full), 'gathered rnd differs'
    d = DistEngine.distances(res, R)
    D = np.zeros((n, n), dtype='<f4')
    iu = np.triu_indices(n, 1)
    D[iu] = d
    D[(iu[1], iu[0])] = d
    assert np.array_equal(D, _D), 'distance matrix differs'
    # matrix path: every rank fills its ranges of the SORTED pair space in a zeroed matrix, one all_reduce merges them
    import torch
    order = np.argsort(-np.array([len(x) for x in seqs]), kind='stable')
    sorted_pairs = [(a, b) for a in range(n) for b in range(a + 1, n)]

    def compute_matrix(shard):
        Dp = torch.zeros((n, n), dtype=torch.float32)
        for (b, e) in (fast_ranges(n_pairs, shard[0], shard[1], min_block=8) if shard else [(0, n_pairs)]):
            for q in range(b, e):
                i, j = order[sorted_pairs[q][0]], order[sorted_pairs[q][1]]
                Dp[i, j] = Dp[j, i] = float(_D[i, j])
        return Dp

    Dm = sharded_dist_matrix(None, seqs, True, True, R, seed, compute=compute_matrix)
    assert Dm.dtype == np.dtype('<f4') and np.array_equal(Dm, _D), 'reduced matrix differs'
    mine = fast_ranges(n_pairs, rank, world, min_block=8)
    assert 0 < sum(e - b for b, e in mine) < n_pairs
    # the product path itself: DistEngine (host orchestration) + sharded_dist_matrix / sharded_dist_matrices over this gloo group,
    # with only the two C entry points answered by the oracle (tests/test_dist_engine_host.py)
    sys.path.insert(0, os.path.join(ROOT, 'tests'))
    from test_dist_engine_host import OracleLib, _engine
    from telogator2_b200 import _lib as tglib, tg_align
    tglib.stream_ptr = lambda t: None
    aligner = tg_align.get_aligner_object(tg_align.get_scoring_matrix('C', 'tvr'), (True, True), 'tvr')
    short = [c[:150] for c in seqs[:9]]
    want_short = oracle.dist_matrix_c(short, P, True, True, R, seed)
    got_short = sharded_dist_matrix(_engine(aligner, OracleLib()), short, True, True, R, seed)
    assert got_short.dtype == np.dtype('<f4') and np.array_equal(got_short, want_short), 'sharded_dist_matrix differs from the oracle'
    # several matrices in one pair space (sharded_dist_matrices) with a stand-in engine that fills its ranges from the oracle
    from telogator2_b200.multi_gpu import sharded_dist_matrices
    groups = [seqs[:5], seqs[5:6], [], seqs[6:14]]
    sizes = np.array([len(g) for g in groups], dtype=np.int64)
    want_many = [oracle.dist_matrix_c(g, P, True, True, R, seed) if len(g) >= 2 else np.zeros((len(g), len(g)), '<f4') for g in groups]

    class FakeEngine:
        stats = {'d2h_bytes': 0}

        def prepare(self, seq_lists, many=False):
            assert many
            return dict(mat_out_off=np.concatenate([[0], np.cumsum(sizes * sizes)]),
                        mat_pair_off=np.concatenate([[0], np.cumsum(sizes * (sizes - 1) // 2)]))

        def matrix_device(self, st, adjust_lens, min_viable, R_, seed_, shard=None):
            flat = torch.zeros(int(st['mat_out_off'][-1]), dtype=torch.float32)
            tot = int(st['mat_pair_off'][-1])
            for (b, e) in (fast_ranges(tot, shard[0], shard[1], min_block=6) if shard else [(0, tot)]):
                for q in range(b, e):
                    m = int(np.searchsorted(st['mat_pair_off'], q, side='right') - 1)
                    nloc = int(sizes[m])
                    pl = [(a, c) for a in range(nloc) for c in range(a + 1, nloc)][q - int(st['mat_pair_off'][m])]
                    for (x, y) in (pl, pl[::-1]):
                        flat[int(st['mat_out_off'][m]) + x * nloc + y] = float(want_many[m][x, y])
            return flat
    got_many = sharded_dist_matrices(FakeEngine(), groups, True, True, R, seed)
    assert len(got_many) == len(groups) and all(np.array_equal(a, b) for a, b in zip(got_many, want_many)), 'batched matrices differ'
    small = [[c[:120] for c in g] for g in groups]
    want_small = [oracle.dist_matrix_c(g, P, True, True, R, seed) if len(g) >= 2 else np.zeros((len(g), len(g)), '<f4') for g in small]
    got_small = sharded_dist_matrices(_engine(aligner, OracleLib()), small, True, True, R, seed)
    assert all(np.array_equal(a, b) for a, b in zip(got_small, want_small)), 'sharded_dist_matrices (real engine) differs'
    # scan: contiguous read shares by base count, per-read results exchanged with all_gather_object (no sequences travel)
    sys.path.insert(0, os.path.join(ROOT, 'tests'))
    from test_host_logic import _fake_tel_local
    from telogator2_b200.multi_gpu import attach_tel_results, shard_reads_by_bases, strip_tel_results
    rng = np.random.default_rng(8)
    reads = [(f'read{i} x', ''.join(rng.choice(list('ACGT'), int(rng.integers(30, 400)))), None) for i in range(41)]
    bnd = shard_reads_by_bases([len(r[1]) for r in reads], world)
    mine_res, mine_meta = _fake_tel_local(reads[bnd[rank]:bnd[rank + 1]], bnd[rank])
    parts = [None] * world
    dist.all_gather_object(parts, strip_tel_results(mine_res, mine_meta))
    assert attach_tel_results(parts, reads) == _fake_tel_local(reads, 0)[0], 'merged scan result differs'
    # the product entry point itself, with the per-rank device work replaced by the stand-in
    from telogator2_b200 import tg_tel
    tg_tel._tel_repeat_comp_local = lambda rd, params, gtt=None, off=0, return_meta=False: (
        _fake_tel_local(rd, off) if return_meta else _fake_tel_local(rd, off)[0])
    assert tg_tel.get_tel_repeat_comp_parallel(reads, None) == _fake_tel_local(reads, 0)[0], 'sharded entry point differs'
    own = shard_pair_indices(n_pairs, rank, world, block=8)
    np.save(out_path + f'.rank{rank}.npy', own)
    dist.barrier()
    dist.destroy_process_group()
    print('rank', rank, 'ok', len(own), 'pairs')


if __name__ == '__main__':
    main()
