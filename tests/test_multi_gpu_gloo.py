"""CPU: the N>1 path (pair sharding + result gather) under a world_size-2 gloo process group."""
import os
import socket
import subprocess
import sys

import numpy as np

from conftest import ROOT


def _free_port():
    s = socket.socket()
    s.bind(('127.0.0.1', 0))
    p = s.getsockname()[1]
    s.close()
    return p


def test_sharded_pair_scores_world2_gloo(tmp_path):
    port = _free_port()
    procs = []
    for rank in range(2):
        env = dict(os.environ, RANK=str(rank), WORLD_SIZE='2', LOCAL_RANK=str(rank), MASTER_ADDR='127.0.0.1',
                   MASTER_PORT=str(port), OMP_NUM_THREADS='2')
        procs.append(subprocess.Popen([sys.executable, os.path.join(ROOT, 'tests', '_gloo_worker.py'), str(tmp_path / 'own')],
                                      env=env, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True))
    outs = [p.communicate(timeout=300)[0] for p in procs]
    for p, o in zip(procs, outs):
        assert p.returncode == 0, o
    own = [np.load(str(tmp_path / 'own') + f'.rank{r}.npy') for r in range(2)]
    assert len(np.intersect1d(own[0], own[1])) == 0 and len(own[0]) + len(own[1]) == 14 * 13 // 2
    assert abs(len(own[0]) - len(own[1])) <= 8
