"""HOST helpers of the C ABI (no GPU): f4l_host_pack_corr_targets and f4l_host_expand_sparse, the two functions the
host-buffer pipeline runs on CPU threads around the PCIe copies."""
import numpy as np
import torch

from fusion4landslide_b200 import ops


def test_host_pack_corr_targets_matches_numpy():
    rng = np.random.default_rng(0)
    n = 200_000                                          # above the single-thread cut-off of the helper
    corr = np.stack([np.arange(n), rng.integers(-1, 500_000, size=n)], axis=1).astype(np.int64)
    corr[::1000, 1] = 2 ** 35                            # out of the int32 range -> "no correspondence"
    corr[5::1000, 1] = -12345
    want = corr[:, 1].copy()
    want[(want < 0) | (want > 2 ** 31 - 1)] = -1
    for threads in (1, 3):
        out = ops.host_pack_corr_targets(torch.from_numpy(corr), n_threads=threads)
        np.testing.assert_array_equal(out.numpy(), want.astype(np.int32))
    # a caller-provided (larger) buffer is filled in place and a view of the first n elements is returned
    buf = torch.full((n + 7,), 99, dtype=torch.int32)
    out = ops.host_pack_corr_targets(torch.from_numpy(corr), buf, 2)
    assert out.data_ptr() == buf.data_ptr() and out.numel() == n and bool((buf[n:] == 99).all())


def test_host_expand_sparse_doubles_every_pair():
    rng = np.random.default_rng(1)
    pair_rows = rng.integers(0, 40, size=300).astype(np.int32)
    pair_rows[::7] = 0
    total = int(pair_rows.sum())
    once = rng.standard_normal((total, 6)).astype(np.float32)
    want, o = [], 0
    for r in pair_rows:
        want += [once[o:o + r], once[o:o + r]]           # base.py:3430,3436: the rows of a pair are appended twice
        o += r
    want = np.concatenate(want)
    for threads in (1, 4):
        out = torch.zeros((2 * total + 3, 6))
        n = ops.host_expand_sparse(torch.from_numpy(once), torch.from_numpy(pair_rows), out, threads)
        assert n == 2 * total
        np.testing.assert_array_equal(out[:n].numpy(), want)
        assert bool((out[n:] == 0).all())
