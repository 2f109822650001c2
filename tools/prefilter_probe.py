"""Timing probe of prefilter_count_kernel: uniform random reads, with/without telomeric reads, two sizes."""
import sys
import numpy as np
import torch
sys.path.insert(0, '.')
import bench
from telogator2_b200 import tg_prefilter, _lib

tail = _lib.load().tg_prefilter_tail_bytes()
_cv, reads, _t = bench.load_fixtures()
tel_src = bench.make_scan_workload(reads, 20_000_000)
g = torch.Generator(device='cuda'); g.manual_seed(1)
for (n, ln, ntel) in ((32768, 16384, 0), (32768, 16384, 512), (131072, 16384, 0), (16384, 131072, 0), (262144, 4096, 0)):
    d = torch.tensor([65, 67, 71, 84], dtype=torch.uint8, device='cuda')[torch.randint(0, 4, (n * ln + tail,), generator=g, device='cuda')]
    lens = np.full(n, ln, dtype=np.int32)
    tel = [r for r in tel_src if len(r) <= ln][:ntel]
    for k, r in enumerate(tel):
        o = (n // max(ntel, 1)) * k * ln
        d[o:o + len(r)] = torch.frombuffer(bytearray(r.encode()), dtype=torch.uint8).cuda()
        lens[(n // ntel) * k] = len(r)
    ar = tg_prefilter.AsciiReads.from_device(d, np.arange(n + 1, dtype=np.int64) * ln, lens)
    ar.count_pair_device('CCCTAACCCTAA', 'TTAGGGTTAGGG'); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ts = []
    for _ in range(5):
        e0.record(); ar.count_pair_device('CCCTAACCCTAA', 'TTAGGGTTAGGG'); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    b = float(lens.astype(np.int64).sum())
    print(f'n={n} len={ln} tel={len(tel)}: {min(ts):.3f} ms  {b / min(ts) / 1e6:.0f} GB/s')
    del d, ar
