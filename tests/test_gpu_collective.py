"""b200s_p2p_reduce_segments (csrc/collective.cu) on ONE GPU: the "peers" are `world` separate buffers of the same device,
so the kernel's whole contract -- segment table, vector offsets, ring order of the sum starting at the caller's own replica,
nothing written outside the segments, every specialised world size and the run-time one -- is checked without a second GPU.
(The multi-GPU use, over torch symmetric memory, is checked against single-GPU gradients by tools/dist_check.py.)"""
import ctypes as C

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _call(bufs, rank, out, segs):
    from my_depthsplat_b200 import _lib
    L = _lib.load()
    world = len(bufs)
    peers = (C.c_void_p * world)(*[b.data_ptr() for b in bufs])
    so = (C.c_ulonglong * len(segs))(*[s[0] for s in segs])
    sn = (C.c_ulonglong * len(segs))(*[s[1] for s in segs])
    return L.b200s_p2p_reduce_segments(peers, world, rank, out.data_ptr(), so, sn, len(segs), torch.cuda.current_stream().cuda_stream)


@pytest.mark.parametrize("world", [2, 3, 4, 8, 16])
def test_p2p_reduce_segments_sums_the_replicas_in_ring_order(world):
    from my_depthsplat_b200 import _lib
    g = torch.Generator().manual_seed(world)
    n = 1 << 16
    bufs = [(torch.randn(n, generator=g) * 10 ** torch.randint(-3, 4, (n,), generator=g).float()).cuda() for _ in range(world)]
    # ragged segment table: tiny, odd multiples of four, one large, zero-length entries skipped
    segs = [(0, 4), (64, 0), (128, 1028), (4096, 40000), (50000, 12), (65532, 4)]
    for rank in (0, world - 1, world // 2):
        out = torch.full((n,), -7.0, device="cuda")
        assert _call(bufs, rank, out, segs) == _lib.B200S_OK
        torch.cuda.synchronize()
        want = np.full(n, -7.0, dtype=np.float32)
        host = [b.cpu().numpy() for b in bufs]
        for off, cnt in segs:
            acc = host[rank][off:off + cnt].copy()
            for k in range(1, world):          # own replica first, then ranks rank+1, rank+2, ... : float32 adds in that order
                acc = acc + host[(rank + k) % world][off:off + cnt]
            want[off:off + cnt] = acc
        np.testing.assert_array_equal(out.cpu().numpy(), want)   # bit for bit, and nothing outside the segments touched


def test_p2p_reduce_segments_rejects_bad_arguments():
    from my_depthsplat_b200 import _lib
    bufs = [torch.zeros(64, device="cuda") for _ in range(2)]
    out = torch.zeros(64, device="cuda")
    assert _call(bufs, 0, out, [(2, 8)]) == _lib.B200S_EBADARG          # offsets and counts are multiples of four floats
    assert _call(bufs, 2, out, [(0, 8)]) == _lib.B200S_EBADARG          # rank outside the world
    assert _call(bufs[:1], 0, out, [(0, 8)]) == _lib.B200S_EBADARG      # a world of one has nothing to pull
    assert _call(bufs, 0, out[1:], [(0, 8)]) == _lib.B200S_EBADARG      # 16-byte alignment of the bases
    torch.cuda.synchronize()
    assert float(out.abs().max()) == 0.0
