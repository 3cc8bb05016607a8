"""CPU: property tests that pin the integer parts of the oracle against independent, naive restatements of the
reference's loops (hypothesis): LengthRegulator, dur_to_mel2ph, monotonic alignment search."""
import itertools

import numpy as np
import torch
from hypothesis import given, settings, strategies as st

from oracle import ctts_oracle as O


@settings(max_examples=60, deadline=None)
@given(st.lists(st.lists(st.integers(min_value=-2, max_value=6), min_size=5, max_size=5), min_size=1, max_size=4),
       st.sampled_from([None, 3, 40]))
def test_length_regulate_matches_the_reference_loop(durs, max_len):
    """modules.py:1222-1249: every row repeated max(int(d), 0) times, zero padded / cropped to max_len."""
    d = torch.tensor(durs, dtype=torch.float32) * 0.5 + 0.25       # fractional: int() truncates toward zero
    B, S = d.shape
    x = torch.arange(B * S * 3, dtype=torch.float32).view(B, S, 3) + 1
    rows, lens = [], []
    for b in range(B):
        r = [x[b, j] for j in range(S) for _ in range(max(int(d[b, j].item()), 0))]
        rows.append(r)
        lens.append(len(r))
    L = max_len if max_len else max(lens)
    if L == 0:
        return
    want = torch.zeros(B, L, 3)
    for b, r in enumerate(rows):
        for t, v in enumerate(r[:L]):
            want[b, t] = v
    got, got_len = O.length_regulate(x, d, max_len)
    assert got_len.tolist() == lens
    assert torch.equal(got, want)


@settings(max_examples=60, deadline=None)
@given(st.lists(st.integers(min_value=0, max_value=5), min_size=1, max_size=8), st.integers(min_value=0, max_value=3))
def test_mel2ph_matches_its_definition(dur, n_pad):
    """utils/tools.py:598-628: frame t belongs to the phoneme whose cumulative duration interval contains t."""
    d = torch.tensor([dur + [3] * n_pad])
    pad = torch.tensor([[False] * len(dur) + [True] * n_pad])
    want = [j + 1 for j, n in enumerate(dur) for _ in range(n)]
    got = O.durations_to_mel2ph(d, pad)[0].tolist()
    assert got == want
    # and it is the inverse of mel2ph_to_dur (utils/tools.py:631-637) on the valid part
    back = np.bincount(np.asarray(got, dtype=np.int64), minlength=len(dur) + 1)[1:].tolist() if got else [0] * len(dur)
    assert back == dur


def _all_monotonic_paths(M, S):
    """Every path that starts in column 0, ends in column S-1 and moves right by 0 or 1 per row."""
    for steps in itertools.product((0, 1), repeat=M - 1):
        if sum(steps) == S - 1:
            cols = [0]
            for s_ in steps:
                cols.append(cols[-1] + s_)
            yield cols


@settings(max_examples=40, deadline=None)
@given(st.integers(min_value=2, max_value=7), st.integers(min_value=1, max_value=4), st.integers(min_value=0, max_value=10**6))
def test_mas_finds_the_best_monotonic_path(M, S, seed):
    """modules.py:36-64: the Viterbi path has the maximal sum of log-probabilities among all monotonic paths."""
    if S > M:
        return
    rng = np.random.default_rng(seed)
    a = rng.random((M, S)).astype(np.float32) + 1e-3
    opt = O.mas_width1(a.copy())
    assert opt.sum() == M and (opt.sum(1) == 1).all()
    cols = opt.argmax(1)
    assert cols[0] == 0 and cols[-1] == S - 1 and (np.diff(cols) >= 0).all() and (np.diff(cols) <= 1).all()
    score = np.log(a)[np.arange(M), cols].sum()
    best = max(np.log(a)[np.arange(M), p].sum() for p in _all_monotonic_paths(M, S))
    assert score >= best - 1e-4


@settings(max_examples=25, deadline=None)
@given(st.integers(min_value=2, max_value=5), st.integers(min_value=1, max_value=6), st.integers(min_value=2, max_value=9),
       st.booleans(), st.integers(min_value=0, max_value=10**6))
def test_batch_norm_train_is_torch_batchnorm_in_training_mode(n, c, t, two_d, seed):
    """_batch_norm_train (training-mode oracle): output AND the buffers left behind equal nn.BatchNorm1d / 2d .train()."""
    g = torch.Generator().manual_seed(seed)
    shape = (n, c, t, 3) if two_d else (n, c, t)
    x = torch.randn(*shape, generator=g) * 2 + 0.5
    bn = (torch.nn.BatchNorm2d if two_d else torch.nn.BatchNorm1d)(c)
    with torch.no_grad():
        bn.weight.copy_(torch.rand(c, generator=g) + 0.5)
        bn.bias.copy_(torch.randn(c, generator=g))
        bn.running_mean.copy_(torch.randn(c, generator=g))
        bn.running_var.copy_(torch.rand(c, generator=g) + 0.5)
    P = {"bn." + k: v.clone() for k, v in bn.state_dict().items()}
    stats = {}
    got = O._batch_norm_train(x, P, "bn.", stats)
    want = bn.train()(x)
    assert torch.allclose(got, want, atol=1e-6, rtol=1e-5)
    after = bn.state_dict()
    for k in ("running_mean", "running_var"):
        assert torch.allclose(stats["bn." + k], after[k], atol=1e-6, rtol=1e-5), k
    assert int(stats["bn.num_batches_tracked"]) == int(after["num_batches_tracked"]) == 1


def test_coord_channels_follow_the_reference_convention():
    """coordconv.py:36-71 with with_r: [input, row coordinate, column coordinate, radius from (0.5, 0.5)], coordinates in
    [-1, 1] (first / last row and column exactly -1 / +1)."""
    x = torch.zeros(2, 1, 5, 4)
    y = O._add_coords_2d(x)
    assert y.shape == (2, 4, 5, 4)
    assert torch.equal(y[:, 1, 0], torch.full((2, 4), -1.0)) and torch.equal(y[:, 1, -1], torch.full((2, 4), 1.0))
    assert torch.equal(y[:, 2, :, 0], torch.full((2, 5), -1.0)) and torch.equal(y[:, 2, :, -1], torch.full((2, 5), 1.0))
    assert torch.allclose(y[0, 3, 0, 0], torch.tensor((1.5 ** 2 * 2) ** 0.5))
