#!/usr/bin/env python3
"""Run a few forward NTTs of 2^LOG points -- target for ncu captures."""
import os, sys
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import numpy as np, torch
from hekaton_system_b200.backend import Context
lg = int(sys.argv[1]) if len(sys.argv) > 1 else 20
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
K = int(sys.argv[3]) if len(sys.argv) > 3 else 1
stream = torch.cuda.Stream(); torch.cuda.set_stream(stream)
ctx = Context("bls12_381", 0, stream=stream.cuda_stream)
a = np.random.default_rng(2).integers(0, 1 << 60, size=(K << lg, 4), dtype=np.uint64)
d = torch.from_numpy(a.view(np.int64)).cuda()
for _ in range(reps):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream); ctx.ntt(d.data_ptr(), lg, K=K); e1.record(stream); torch.cuda.synchronize()
    print(f"ntt 2^{lg} K={K}: {e0.elapsed_time(e1):.4f} ms")
ctx.close()
