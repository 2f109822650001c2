"""Run under torchrun on N GPUs: the sharded tg_tel.get_tel_repeat_comp_parallel must return, on every rank, exactly what
the single-GPU path returns for the whole read list (scan shards by reads, results exchanged at the end)."""
import gzip
import json
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'tests'))
from test_gpu_tel_batch import make_gtt  # noqa: E402
from telogator2_b200 import tg_kmer, tg_tel  # noqa: E402


def main():
    rank, local = int(os.environ['RANK']), int(os.environ['LOCAL_RANK'])
    torch.cuda.set_device(local)
    dist.init_process_group('nccl', device_id=torch.device('cuda', local))
    G = os.path.join(ROOT, 'tests', 'golden')
    T = json.load(open(os.path.join(G, 'kmer_tables.json')))['human_hifi']
    with gzip.open(os.path.join(G, 'reads.json.gz'), 'rt') as f:
        R = json.load(f)
    KL, ISSUB, CANON = T['KMER_LIST'], T['KMER_ISSUBSTRING'], T['CANONICAL_STRINGS']
    KLR, CANON_REV = [tg_kmer.RC(k) for k in KL], [tg_kmer.RC(k) for k in CANON]
    reads = [(name + ' extra', seq, None) for name, seq in R['human_hifi'] + R['human_ont']]
    reads += [(n, s, None) for n, s in R['synthetic'] if len(s) > 0 and set(s) <= set('ACGTN')]
    reads = [(f'{n}#{k}', s, q) for k in range(3) for (n, s, q) in reads]
    params = [T['read_type'], 60, KL, KLR, ISSUB, CANON, CANON_REV, 100, 0.5, 100, 400, 100, 1000, False, '', 100, 1000]
    for gtt in (None, make_gtt(tg_kmer.get_telomere_kmer_density)):       # boundary calling on the device / through a callback
        got = tg_tel.get_tel_repeat_comp_parallel(reads, params, get_terminating_tl=gtt)
        want = tg_tel._tel_repeat_comp_local(reads, params, gtt, 0)
        assert got == want, 'sharded scan result differs from the single-GPU result'
        assert len(got[0]) > 0
    dist.barrier()
    print(f'rank {rank}: sharded scan == local scan ({len(reads)} reads, {len(got[0])} tvr entries, {len(got[6])} interstitial)')
    dist.destroy_process_group()


if __name__ == '__main__':
    main()
