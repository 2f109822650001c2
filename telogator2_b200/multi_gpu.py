"""Sharding of the pair space over the ranks of one box (torch.distributed, one process per GPU).

Pairs are independent and a pair's RNG seed is a pure function of its global index
(rng_seed + p, tg_align.py:157,170-172), so sharding cannot change any result.  Blocks of
consecutive pair indices are dealt round-robin to the ranks (statistical cost balance); the only
exchange step is the gather of the per-pair integer results over NCCL/NVLink: an all_gather of each rank's compact
share (device-driven path) or element-wise all-reduces of full-length partial arrays (explicit-job path)."""
import numpy as np

SHARD_BLOCK = 4096


def shard_pair_indices(n_pairs, rank, world, block=SHARD_BLOCK):
    """Sorted global pair indices owned by `rank`."""
    if world <= 1:
        return np.arange(n_pairs, dtype=np.int64)
    starts = np.arange(rank * block, n_pairs, world * block, dtype=np.int64)
    if len(starts) == 0:
        return np.zeros(0, dtype=np.int64)
    idx = (starts[:, None] + np.arange(block, dtype=np.int64)[None, :]).reshape(-1)
    return idx[idx < n_pairs]


_LOCAL_ONLY = 0


class local_only:
    """Context manager: inside it the sharding entry points behave as in a single-process run even if a process group is
    initialised -- for code that only SOME ranks execute (a collective there would hang the job)."""

    def __enter__(self):
        global _LOCAL_ONLY
        _LOCAL_ONLY += 1
        return self

    def __exit__(self, *exc):
        global _LOCAL_ONLY
        _LOCAL_ONLY -= 1
        return False


def _dist_state():
    if _LOCAL_ONLY:
        return None, 0, 1
    try:
        import torch.distributed as dist
    except Exception:                                   # noqa: BLE001
        return None, 0, 1
    if not (dist.is_available() and dist.is_initialized()):
        return None, 0, 1
    return dist, dist.get_rank(), dist.get_world_size()


def gather_results(res, dist, device):
    """Merge every rank's partial result dict into full-length arrays on every rank.
    Compact results of the device-driven path (res['ranges']) are exchanged with all_gather (each rank's share,
    padded to the longest); full-length partial results are merged with element-wise all-reduces."""
    import torch
    if 'ranges' in res:
        from .dist_engine import NOT_COMPUTED, fast_ranges
        n_pairs = res['n_pairs']
        world = dist.get_world_size()
        all_ranges = [fast_ranges(n_pairs, r, world) for r in range(world)]
        n_max = max(sum(e - b for b, e in rr) for rr in all_ranges)
        out = {k: v for k, v in res.items() if k not in ('ranges', 'n_pairs')}
        for key, fill in (('aln', NOT_COMPUTED), ('rnd', NOT_COMPUTED), ('iden', 0), ('n', 0), ('m', 0)):
            src = np.ascontiguousarray(res[key])
            if src.size == 0 and src.ndim > 1:
                out[key] = np.full((n_pairs,) + src.shape[1:], fill, dtype=src.dtype)
                continue
            pad = np.full((n_max,) + src.shape[1:], fill, dtype=src.dtype)
            pad[:len(src)] = src
            mine = torch.from_numpy(pad).to(device)
            parts = [torch.empty_like(mine) for _ in range(world)]
            dist.all_gather(parts, mine)
            full = np.full((n_pairs,) + src.shape[1:], fill, dtype=src.dtype)
            for r in range(world):
                part = parts[r].cpu().numpy()
                pos = 0
                for (b, e) in all_ranges[r]:
                    full[b:e] = part[pos:pos + (e - b)]
                    pos += e - b
            out[key] = full
        return out
    for key, op in (('aln', dist.ReduceOp.MAX), ('rnd', dist.ReduceOp.MAX), ('n', dist.ReduceOp.MAX),
                    ('m', dist.ReduceOp.MAX), ('iden', dist.ReduceOp.SUM)):
        arr = res[key]
        if arr.size == 0:
            continue
        t = torch.from_numpy(np.ascontiguousarray(arr)).to(device)
        dist.all_reduce(t, op=op)
        res[key] = t.cpu().numpy().reshape(arr.shape)
    return res


def sharded_pair_scores(engine, sequences, adjust_lens, min_viable, R, rng_seed, compute=None, device=None):
    """engine.pair_scores over this rank's shard + gather.  `compute(shard)` with shard = (rank, world) replaces the device
    call in the CPU (gloo) tests of the sharding logic."""
    dist, rank, world = _dist_state()
    n = len(sequences)
    n_pairs = n * (n - 1) // 2
    if compute is None:
        def compute(shard):
            return engine.pair_scores(sequences, adjust_lens, min_viable, R, rng_seed, shard=shard)
    if world <= 1 or n_pairs == 0:
        return compute(None)
    res = compute((rank, world))
    if device is None:
        device = engine.device if dist.get_backend() == 'nccl' else 'cpu'
    return gather_results(res, dist, device)


def _to_host(D):
    """Device tensor -> ndarray through a page-locked block (torch caches and reuses these): the copy runs at PCIe speed and
    the returned ndarray owns the block until it is garbage collected.  CPU tensors (gloo tests) are returned as they are."""
    if not D.is_cuda:
        return D.numpy()
    import torch
    host = torch.empty(D.shape, dtype=D.dtype, pin_memory=True)
    host.copy_(D)
    return host.numpy()


def reduce_matrix(D, dist):
    """Sum the ranks' partial matrices (disjoint non-zero entries) in place: the one exchange step of the matrix path."""
    dist.all_reduce(D, op=dist.ReduceOp.SUM)
    return D


def sharded_dist_matrix(engine, sequences, adjust_lens, min_viable, R, rng_seed, compute=None):
    """get_dist_matrix_parallel (tg_align.py:154-189) over every rank of the process group: each rank fills the entries of
    its share of the pair space in a zeroed float32 matrix on its device, the matrices are summed over NCCL/NVLink
    (x + 0 = x exactly, NaN stays NaN), and every rank returns the full host matrix.
    `compute(shard)` -> partial matrix (torch tensor) replaces the device call in the CPU (gloo) tests."""
    dist, rank, world = _dist_state()
    n = len(sequences)
    if n < 2:
        return np.zeros((n, n), dtype='<f4')
    if compute is None:
        st = engine.prepare(sequences)
        if st is None:                       # scoring/length outside the device-driven path: integer gather + host epilogue
            res = sharded_pair_scores(engine, sequences, adjust_lens, min_viable, R, rng_seed)
            return engine.dist_matrix(sequences, adjust_lens, min_viable, R, rng_seed, res=res)

        def compute(shard):
            return engine.matrix_device(st, adjust_lens, min_viable, R, rng_seed, shard=shard)
    D = compute((rank, world) if world > 1 else None)
    if world > 1:
        D = reduce_matrix(D, dist)
    out = _to_host(D)
    if engine is not None:
        engine.stats['d2h_bytes'] += out.nbytes
    return out


def sharded_dist_matrices(engine, seq_lists, adjust_lens, min_viable, R, rng_seed):
    """Several independent get_dist_matrix_parallel calls (one per entry of seq_lists, all with the same aligner and
    rng_seed) as ONE device pass: the matrices' pair spaces are concatenated, so small per-cluster matrices (iterations 1/2,
    sub-telomere stage) fill the GPU together, and under a process group the concatenated space is what gets
    sharded.  Returns a list of float32 (n_b, n_b) host matrices, identical to the per-call results."""
    dist, rank, world = _dist_state()
    seq_lists = [list(g) for g in seq_lists]
    st = engine.prepare(seq_lists, many=True) if any(len(g) >= 2 for g in seq_lists) else None
    if st is None:                            # nothing to align, or outside the device-driven path: one call per matrix
        if any(len(g) >= 2 for g in seq_lists):
            return [sharded_dist_matrix(engine, g, adjust_lens, min_viable, R, rng_seed) for g in seq_lists]
        return [np.zeros((len(g), len(g)), dtype='<f4') for g in seq_lists]
    D = engine.matrix_device(st, adjust_lens, min_viable, R, rng_seed, shard=(rank, world) if world > 1 else None)
    if world > 1:
        D = reduce_matrix(D, dist)
    flat = _to_host(D)
    engine.stats['d2h_bytes'] += flat.nbytes
    off = st['mat_out_off']
    return [flat[off[b]:off[b + 1]].reshape(len(g), len(g)) for b, g in enumerate(seq_lists)]


# ---------------------------------------------------------------------------------------------------------
# scan: reads are independent -> contiguous shares of equal base count, no exchange step on the data path; the per-read
# results are exchanged once at the end
# ---------------------------------------------------------------------------------------------------------
def shard_reads_by_bases(lens, world):
    """world + 1 boundaries: rank r owns reads [b[r], b[r+1]) -- contiguous, cumulative base count as even as possible."""
    lens = np.asarray(lens, dtype=np.int64)
    cum = np.concatenate([[0], np.cumsum(lens)])
    targets = cum[-1] * np.arange(1, world, dtype=np.float64) / world
    inner = np.searchsorted(cum, targets, side='left')
    b = np.concatenate([[0], inner, [len(lens)]]).astype(np.int64)
    return [int(x) for x in np.maximum.accumulate(b)]


def strip_tel_results(res, meta):
    """A rank's result list with every read sequence replaced by (global read index, reverse-complemented?) (from
    _tel_repeat_comp_local's meta), so that only hits and numbers travel.  Entries are [hits, tlen, 0, 'q', name, mapq, read];
    the interstitial sub-telomere pairs [(name_left, left), (name_right, right)] keep their names and lengths."""
    out = list(res)
    out[0] = [e[:6] + [m] for e, m in zip(res[0], meta['tvr'])]
    keys = list(res[7].keys())
    if len(keys) == len(meta['interstitial']) == len(res[6]):
        out[6] = [[(nm, len(seq)) for (nm, seq) in pair] for pair in res[6]]
        out[7] = {rn: res[7][rn][:6] + [m] for rn, m in zip(keys, meta['interstitial'])}
        out[8] = {rn: res[8][rn][:6] + [m] for rn, m in zip(keys, meta['interstitial'])}
        out.append(True)
    else:                       # duplicate read names collapsed dictionary entries (as upstream): ship that part as it is
        out.append(False)
    return out


def attach_tel_results(parts, all_read_dat):
    """Merge the stripped per-rank results (rank order = read order) and put the sequences back."""
    from .tg_util import RC

    def attach(e):
        idx, flipped = e[6]
        seq = all_read_dat[idx][1]
        return e[:6] + [RC(seq) if flipped else seq]
    out = [[], [], [], 0, 0, 0, [], {}, {}]
    for p in parts:
        out[0] += [attach(e) for e in p[0]]
        out[1] += p[1]
        out[2] += p[2]
        out[3] += p[3]
        out[4] += p[4]
        out[5] += p[5]
        if not p[9]:
            out[6] += p[6]
            out[7].update(p[7])
            out[8].update(p[8])
            continue
        for pair, rn in zip(p[6], p[7].keys()):
            out[7][rn] = attach(p[7][rn])
            out[8][rn] = attach(p[8][rn])
            # interstitial sub-telomeres (tg_tel.py:283-286): left = read[:len_left], right = read[len - len_right:]
            seq = out[7][rn][6]
            (nl, ll), (nr, lr) = pair
            out[6].append([(nl, seq[:ll]), (nr, seq[len(seq) - lr:])])
    return out
