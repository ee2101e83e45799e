"""CPU tests for the hypercolumn sampler: the oracle restatement of
Sampling._sample (nn/strotss_utils.py:25-81) and the host-side index generator mirror (:83-121)."""
import numpy as np
import pytest
import torch

from oracle import strotss_oracle as O

SHAPES = [(341, 512, 3), (341, 512, 64), (341, 512, 64), (170, 256, 128), (170, 256, 128), (85, 128, 256),
          (85, 128, 256), (85, 128, 256), (42, 64, 512), (42, 64, 512)]           # content at scale 512


def test_scale_axis_rule_of_the_reference():
    # 170 is not a power of two -> the WIDTH ratio is used for every level (strotss_utils.py:35-36)
    assert O.sampler_scales(SHAPES) == [1.0, 1.0, 1.0, 2.0, 1.0, 2.0, 1.0, 1.0, 2.0, 1.0]
    # power-of-two heights -> the HEIGHT ratio
    sq = [(64, 48, 3), (64, 48, 8), (32, 24, 8), (16, 12, 8)]
    assert O.sampler_scales(sq) == [1.0, 1.0, 2.0, 2.0]
    odd = [(85, 128, 3), (42, 64, 8)]
    assert O.sampler_scales(odd)[1] == 2.0 and O.sampler_scales([(85, 127, 3), (42, 63, 8)])[1] == pytest.approx(127 / 63)


def test_total_width_is_2179_and_integer_indices_make_bilinear_equal_nearest():
    rng = np.random.default_rng(0)
    small = [(20, 24, 3), (20, 24, 5), (10, 12, 7)]
    xs = [rng.standard_normal((1,) + s).astype(np.float32) for s in small]
    idx = np.stack([rng.integers(0, 20, 30), rng.integers(0, 24, 30)], axis=1).astype(np.float32)
    idx = (idx // 2) * 2                       # stay integral after the /2 of the last level
    a = O.sample_hypercolumns(xs, idx, True)
    b = O.sample_hypercolumns(xs, idx, False)
    assert a.shape == (30, 15) and np.array_equal(a, b)
    assert sum(s[2] for s in SHAPES) == 2179


def test_bilinear_weights_and_border_clipping():
    x = np.arange(12, dtype=np.float32).reshape(1, 3, 4, 1)
    out = O.sample_hypercolumns([x], np.array([[0.5, 1.25]], dtype=np.float32), True)
    # rows 0/1, cols 1/2: values 1,2,5,6
    assert out[0, 0] == pytest.approx(0.5 * 0.75 * 1 + 0.5 * 0.25 * 2 + 0.5 * 0.75 * 5 + 0.5 * 0.25 * 6)
    edge = O.sample_hypercolumns([x], np.array([[2.6, 3.9]], dtype=np.float32), True)
    assert edge[0, 0] == pytest.approx(11.0)   # both +1 taps clip back onto the last row / column


def test_index_generator_mirror():
    from strotss_tensorflow_b200.sampling import Sampling
    s = Sampling(1024, torch.Generator().manual_seed(0))
    base = torch.zeros(1, 341, 512, 3)
    idx = s._make_indices(base, True)
    assert idx.shape == (1024, 2) and idx.dtype == torch.float32
    assert float(idx[:, 0].max()) < 341 and float(idx[:, 1].max()) < 512
    assert len({(int(a), int(b)) for a, b in idx.tolist()}) == 1024            # a sample without replacement
    # strided grid: step (3, 4) at 341x512 -> all rows congruent mod 3, all columns congruent mod 4
    grid, steps = O.sampler_grid(341, 512, True)
    assert steps == (3, 4)
    assert len(set((idx[:, 0] % 3).tolist())) == 1 and len(set((idx[:, 1] % 4).tolist())) == 1
    # mask: only positions inside the region; an all-zero mask falls back to everything (:107-108)
    m = torch.zeros(341, 512, 1); m[:100, :200] = 1
    im = s._make_indices(base, True, m)
    assert float(im[:, 0].max()) < 100 and float(im[:, 1].max()) < 200
    assert s._make_indices(base, True, torch.zeros(341, 512, 1)).shape == (1024, 2)
    # a region with fewer candidates than sample_size yields a ragged N_r (masked mode, :113,120)
    tiny = torch.zeros(341, 512, 1); tiny[:30, :40] = 1
    assert 0 < s._make_indices(base, True, tiny).shape[0] < 1024
