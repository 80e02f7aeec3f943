"""Micro-benchmarks of single stages through the C ABI (GPU box only).  python tools/microbench.py sort [n] [bits]"""
import sys
from pathlib import Path

import numpy as np
import torch

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from my_depthsplat_b200 import _lib  # noqa: E402


def bench_sort(n=34_000_000, bits=47, iters=4):
    L = _lib.load()
    g = torch.Generator(device="cuda").manual_seed(0)
    depth = (torch.rand(n, device="cuda", generator=g) * 19 + 1).view(torch.int32).to(torch.int64)
    tile = torch.randint(0, 1 << (bits - 32), (n,), device="cuda", generator=g, dtype=torch.int64)
    keys0 = (tile << 32) | depth
    vals0 = torch.arange(n, device="cuda", dtype=torch.int32)
    ka, va = keys0.clone(), vals0.clone()
    kb, vb = torch.empty_like(ka), torch.empty_like(va)
    tmp = torch.empty(L.b200s_sort_tmp_bytes(n), dtype=torch.uint8, device="cuda")
    s = torch.cuda.current_stream().cuda_stream
    ms = []
    for i in range(iters + 2):
        ka.copy_(keys0); va.copy_(vals0)
        _lib.profile_enable(True)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        _lib.check(L.b200s_sort_pairs(ka.data_ptr(), va.data_ptr(), kb.data_ptr(), vb.data_ptr(), n, bits, tmp.data_ptr(), s), "sort")
        e1.record(); torch.cuda.synchronize()
        if i >= 2:
            ms.append(e0.elapsed_time(e1))
    _lib.profile_read()
    ka.copy_(keys0); va.copy_(vals0)
    _lib.check(L.b200s_sort_pairs(ka.data_ptr(), va.data_ptr(), kb.data_ptr(), vb.data_ptr(), n, bits, tmp.data_ptr(), s), "sort")
    torch.cuda.synchronize()
    print({k: round(v, 3) for k, v in _lib.profile_read().items() if v > 0}, end=" ")
    ok = bool((ka[1:] >= ka[:-1]).all())
    passes = (bits + 7) // 8
    t = float(np.median(ms))
    print(f"sort n={n} bits={bits} passes={passes}: {t:.3f} ms  ({n / t / 1e6:.1f} Gkeys/s... {(8 + 24 * passes) * n / t / 1e6:.0f} GB/s alg)  sorted={ok}")


if __name__ == "__main__":
    what = sys.argv[1] if len(sys.argv) > 1 else "sort"
    if what == "sort":
        args = [int(a) for a in sys.argv[2:4]]
        for hv in (2,):
            for rv in (0, 1, 2):
                _lib.load().b200s_debug_set(0, hv); _lib.load().b200s_debug_set(1, rv)
                print(f"hist variant {hv}, rank variant {rv}:", end=" ")
                bench_sort(*args)
