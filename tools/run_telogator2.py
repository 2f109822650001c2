#!/usr/bin/env python
"""Run the reference's telogator2.py (or make_telogator_ref.py) UNEDITED over the B200 hot path:

    TELOGATOR2_REFERENCE=/path/to/telogator2 python tools/run_telogator2.py [telogator2.py] -i reads.fa.gz -o out -r hifi --rng 1 ...

The script is executed with this repository ahead of its own directory on sys.path, so its `from source.X import ...` lines
bind to the overlay package source/ of this repository.  Under `torchrun --nproc-per-node N` the distance matrices and the scan
are sharded over N GPUs (telogator2_b200.multi_gpu)."""
import os
import runpy
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from source._overlay import reference_root  # noqa: E402


def main():
    ref = reference_root()
    args = sys.argv[1:]
    script = 'telogator2.py'
    if args and args[0].endswith('.py'):
        script, args = args[0], args[1:]
    path = script if os.path.isabs(script) else os.path.join(ref, script)
    sys.path.insert(1, ref)                       # after this repository: `source` is the overlay, the rest is the reference's
    sys.argv = [path] + args
    if os.environ.get('RANK') is not None and int(os.environ.get('WORLD_SIZE', '1')) > 1:
        import torch
        import torch.distributed as dist
        torch.cuda.set_device(int(os.environ.get('LOCAL_RANK', '0')))
        dist.init_process_group('nccl')
    runpy.run_path(path, run_name='__main__')


if __name__ == '__main__':
    main()
