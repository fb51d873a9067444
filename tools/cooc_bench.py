"""text8-sized co-occurrence counting (cymf/glove.pyx:218-221): 17 M kept tokens, ~71 k words, window 10 = 170 M map
updates.  Device path (cymf_cooc_count_dev) timed with CUDA events; the CPU oracle restatement timed on a 1 M-token
prefix.  Run on the GPU box: python tools/cooc_bench.py"""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402
import torch  # noqa: E402

from cymf_b200 import _lib  # noqa: E402
from oracle import oracle  # noqa: E402

T, V, W = 17_000_000, 71_000, 10
rng = np.random.default_rng(8)
p = 1.0 / (np.arange(V) + 1.0)
tokens = rng.choice(V, size=T, p=p / p.sum()).astype(np.int32)
pos = np.arange(T, dtype=np.int32)                            # one line, like text8
L = _lib.lib()
d_tok, d_pos = torch.from_numpy(tokens).cuda(), torch.from_numpy(pos).cuda()
cap = T * W
rows = torch.empty(cap, dtype=torch.int32, device="cuda")
cols = torch.empty(cap, dtype=torch.int32, device="cuda")
vals = torch.empty(cap, dtype=torch.float64, device="cuda")
nnz = torch.zeros(1, dtype=torch.int64, device="cuda")
ws = torch.empty(int(L.cymf_cooc_workspace_bytes(T, W)), dtype=torch.uint8, device="cuda")


def run():
    _lib.check(L.cymf_cooc_count_dev(_lib.ptr(d_tok), _lib.ptr(d_pos), T, V, W, _lib.ptr(rows), _lib.ptr(cols),
                                     _lib.ptr(vals), cap, _lib.ptr(nnz), _lib.ptr(ws), _lib.stream_ptr()))


run()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(3):
    run()
e1.record()
torch.cuda.synchronize()
sec = e0.elapsed_time(e1) * 1e-3 / 3
m = int(nnz.item())
n_cpu = 1_000_000
t0 = time.perf_counter()
r, c, v = oracle.cooc_count([tokens[:n_cpu].tolist()], V, W)
cpu = time.perf_counter() - t0
# same prefix on the device: bit-identical cells
d_tok2, d_pos2 = d_tok[:n_cpu].contiguous(), d_pos[:n_cpu].contiguous()
_lib.check(L.cymf_cooc_count_dev(_lib.ptr(d_tok2), _lib.ptr(d_pos2), n_cpu, V, W, _lib.ptr(rows), _lib.ptr(cols),
                                 _lib.ptr(vals), cap, _lib.ptr(nnz), _lib.ptr(ws), _lib.stream_ptr()))
m2 = int(nnz.item())
same = (m2 == r.shape[0] and np.array_equal(rows[:m2].cpu().numpy(), r) and np.array_equal(cols[:m2].cpu().numpy(), c)
        and np.array_equal(vals[:m2].cpu().numpy(), v))
print(json.dumps({"tokens": T, "vocab": V, "window": W, "map_updates": T * W, "cells": m, "device_sec": sec,
                  "updates_per_s": T * W / sec, "workspace_GB": ws.numel() / 1e9,
                  "cpu_oracle": {"tokens": n_cpu, "sec": cpu, "updates_per_s": n_cpu * W / cpu, "cores": 1},
                  "prefix_bit_identical_to_oracle": bool(same)}))
