"""The BINNED sort mode's per-bin segment sort (csrc/binsort.cu) through b200s_segment_sort, against numpy: every bin's
values in ascending (key, value) order -- what a stable sort of pairs emitted in value order produces -- whatever the
arrival order inside the bin.  Covers the three shared-memory size classes of the bucket path, the global-memory LSD path
(bins longer than 26 976 entries, and bins handed over because a bucket was crowded), empty bins, duplicate keys (short
runs: rank inside the bucket; long runs: the LSD path's index passes) and narrow / wide key ranges."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _run(counts, keys, vals):
    from my_depthsplat_b200 import _lib
    L = _lib.load()
    n, bins = int(keys.shape[0]), int(counts.shape[0])
    ent = (vals.astype(np.uint64) << np.uint64(32)) | keys.astype(np.uint64)   # uint2 (x = key, y = value), little endian
    d_ent = torch.from_numpy(ent.view(np.int64)).cuda()
    d_cnt = torch.from_numpy(counts.astype(np.int32)).cuda()
    d_vals = torch.full((max(n, 1),), -1, dtype=torch.int32, device="cuda")
    d_rng = torch.zeros((bins, 2), dtype=torch.int32, device="cuda")
    tmp = torch.zeros(L.b200s_segment_sort_tmp_bytes(n, bins) + 256, dtype=torch.uint8, device="cuda")
    _lib.check(L.b200s_segment_sort(d_cnt.data_ptr(), bins, d_ent.data_ptr(), n, d_vals.data_ptr(), d_rng.data_ptr(), tmp.data_ptr(),
                                    torch.cuda.current_stream().cuda_stream), "segment_sort")
    torch.cuda.synchronize()
    return d_vals.cpu().numpy().view(np.uint32)[:n], d_rng.cpu().numpy().view(np.uint32)


def _expect(counts, keys, vals):
    starts = np.concatenate([[0], np.cumsum(counts)])
    out = np.empty_like(vals)
    for b in range(len(counts)):
        s, e = starts[b], starts[b + 1]
        order = np.lexsort((vals[s:e], keys[s:e]))
        out[s:e] = vals[s:e][order]
    return out, np.stack([starts[:-1], starts[1:]], 1).astype(np.uint32)


CASES = {
    "mixed_classes": dict(counts=[0, 1, 31, 33, 65, 700, 4960, 4961, 12224, 12225, 0, 26976, 26977, 70000, 5], keybits=32, dup=0.0),
    "narrow_range_ties": dict(counts=[4000, 9000, 100, 20000], keybits=9, dup=0.0),       # 512 distinct keys: runs of ~10-40
    "float_depths": dict(counts=[4400] * 40 + [8300] * 8, keybits=None, dup=0.001),
    "all_equal_keys": dict(counts=[3000, 7000, 15000, 64], keybits=0, dup=0.0),          # one run per bin: index-pass fallback
    "many_small": dict(counts=list(np.random.default_rng(5).integers(0, 300, size=3000)), keybits=20, dup=0.01),
}


@pytest.mark.parametrize("name", list(CASES))
def test_segment_sort_matches_numpy(name):
    c = CASES[name]
    rng = np.random.default_rng(abs(hash(name)) % (2 ** 31))
    counts = np.asarray(c["counts"], dtype=np.int64)
    n = int(counts.sum())
    if c["keybits"] is None:   # depth-like floats: clustered surfaces, bits of positive floats
        z = np.where(rng.random(n) < 0.7, 5.0 + 0.05 * rng.standard_normal(n), 2.0 + 18.0 * rng.random(n)).astype(np.float32)
        keys = np.abs(z).view(np.uint32).copy()
    elif c["keybits"] == 0:
        keys = np.full(n, 0x40490FDB, dtype=np.uint32)
    else:
        keys = (rng.integers(0, 2 ** c["keybits"], size=n, dtype=np.uint64) + np.uint64(0x3F000000 if c["keybits"] < 30 else 0)).astype(np.uint32)
    if c["dup"] > 0 and n > 1:
        m = rng.random(n) < c["dup"]
        keys[m] = keys[np.maximum(np.nonzero(m)[0] - 1, 0)]
    vals = rng.permutation(max(n, 1)).astype(np.uint32)[:n] * 3 + 7   # unique indices, arbitrary arrival order
    got_v, got_r = _run(counts, keys, vals)
    want_v, want_r = _expect(counts, keys, vals)
    np.testing.assert_array_equal(got_r, want_r)
    np.testing.assert_array_equal(got_v, want_v)
