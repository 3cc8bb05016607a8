"""GPU: per-kernel parity of the forward kernels that round 1 only covered end to end -- MAS (ties, one-phoneme
utterances, ragged lengths, more phonemes than frames), aligner attention, bidirectional GRU, fastformer pooling,
relative-shift softmax, depthwise conv + BatchNorm + Swish, GLU, phoneme-level energy -- each against its CPU restatement
(oracle/capi_emulator.py, which follows the reference lines cited in include/ctts_b200.h)."""
import os
import sys

import numpy as np
import pytest
import torch

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from gpu_harness import Scratch, g, run_both  # noqa: E402

pytestmark = pytest.mark.gpu


def _soft_attention(B, M, S, src_lens, mel_lens, seed, ties=False):
    gen = torch.Generator().manual_seed(seed)
    a = torch.rand(B, 1, M, S, generator=gen) + 1e-3
    if ties:      # many exactly equal probabilities: the `>=` tie rule of modules.py:51 decides the path
        a = (a * 4).floor() / 4 + 0.25
    for b in range(B):
        a[b, :, :, src_lens[b]:] = 0
        a[b, :, mel_lens[b]:, :] = 0
    return a / a.sum(-1, keepdim=True).clamp(min=1e-8)


@pytest.mark.parametrize("case", ["ragged", "ties", "one_phoneme", "more_phonemes_than_frames", "long"])
def test_mas_matches_reference_semantics(case):
    """ctts_mas against mas_width1 / b_mas (modules.py:36-75) through the oracle: hard path and durations BIT-EXACT."""
    if case == "ragged":
        B, M, S, sl, ml = 4, 37, 11, [11, 7, 3, 1], [37, 30, 9, 5]
    elif case == "ties":
        B, M, S, sl, ml = 3, 25, 8, [8, 8, 5], [25, 20, 25]
    elif case == "one_phoneme":
        B, M, S, sl, ml = 2, 12, 4, [1, 1], [12, 1]
    elif case == "more_phonemes_than_frames":
        B, M, S, sl, ml = 2, 6, 9, [9, 8], [4, 6]
    else:
        B, M, S, sl, ml = 2, 400, 60, [60, 41], [400, 333]
    attn = _soft_attention(B, M, S, sl, ml, seed=len(case), ties=(case == "ties"))
    src_lens, mel_lens = torch.tensor(sl), torch.tensor(ml)
    ws = torch.zeros(B * M * S, dtype=torch.uint8)
    hard, dur = torch.zeros(B, 1, M, S), torch.zeros(B, S)
    run_both("ctts_mas", [attn, src_lens, mel_lens, B, M, S, ws, hard, dur], atol=0, rtol=0, int_exact=False)


def test_aligner_attention():
    B, M, S, C = 3, 45, 19, 80
    lens = torch.tensor([19, 10, 1])
    prior = torch.rand(B, S, M, generator=torch.Generator().manual_seed(1))
    run_both("ctts_aligner_attention", [g(B, M, C) * 3, g(B, S, C, seed=1) * 3, prior, lens, 0.0005, B, M, S, C,
                                        torch.zeros(B, M, S), torch.zeros(B, M, S)], atol=2e-6, rtol=1e-5)


def test_gru_bidir():
    B, T, H = 3, 23, 128
    args = [g(B, T, 3 * H), g(B, T, 3 * H, seed=1), g(3 * H, H, seed=2) * 0.1, g(3 * H, seed=3) * 0.1, g(3 * H, H, seed=4) * 0.1,
            g(3 * H, seed=5) * 0.1, B, T, H, torch.zeros(B, T, 2 * H), torch.zeros(B, 2 * H)]
    run_both("ctts_gru_bidir", args, atol=2e-5, rtol=1e-4)


def test_fastformer_pool_inverted_mask():
    B, T, heads, hs = 3, 50, 128, 2
    lens = torch.tensor([50, 31, 50])        # padded AND unpadded utterances (the inverted mask treats them differently)
    run_both("ctts_fastformer_pool", [g(B, T, heads), g(B, T, heads * hs, seed=1), lens, B, T, heads, hs,
                                      torch.zeros(B, heads * hs)], atol=1e-5, rtol=1e-4)


def test_relshift_softmax():
    Z, T = 5, 37
    ldp = 48
    run_both("ctts_relshift_softmax", [g(Z, T, T), g(Z, T, T, seed=1), Z, T, ldp, 16.0, torch.zeros(Z, T, ldp)], atol=1e-6)


def test_glu_and_dwconv_bn_swish():
    rows, C = 200, 256
    run_both("ctts_glu", [g(rows, 2 * C), rows, C, torch.zeros(rows, C)], atol=1e-6)
    B, T, C, K = 2, 70, 256, 31
    run_both("ctts_dwconv_bn_swish", [g(B, T, C), g(C, K, seed=1) * 0.2, K, 1 + 0.1 * g(C, seed=2), 0.1 * g(C, seed=3), B, T, C,
                                      torch.zeros(B, T, C)], atol=2e-5)


def test_phoneme_energy_sequential_in_place_semantics():
    """get_phoneme_level_energy (utils/tools.py:56-66) overwrites the frame array while it walks it: zero durations and
    durations that overrun M included."""
    B, S, M = 3, 9, 30
    dur = torch.tensor([[3, 0, 5, 1, 1, 7, 2, 0, 4], [1, 1, 1, 1, 1, 1, 1, 1, 1], [10, 10, 10, 0, 0, 0, 0, 0, 0]]).float()
    lens = torch.tensor([9, 9, 3])
    run_both("ctts_phoneme_energy", [dur, lens, g(B, M), B, S, M, Scratch(torch.zeros(B * M)), torch.zeros(B, S)], atol=1e-6)


def test_conformer_tensor_core_attention_helpers():
    from gpu_harness import PA
    Z, T, ld, ldp = 4, 37, 40, 40
    pl = [torch.zeros(Z, T, ldp, dtype=torch.bfloat16) for _ in range(2)]
    run_both("ctts_relshift_softmax_planes", [g(Z, T, ld), g(Z, T, ld, seed=1), Z, T, ld, ldp, 16.0, 2, PA(pl)], atol=2e-5)
    rows, H, DH, DHp = 50, 8, 32, 64
    pl = [torch.zeros(rows, H * DHp, dtype=torch.bfloat16) for _ in range(3)]
    run_both("ctts_pad_heads_planes", [g(rows, 3 * H * DH), g(H * DH, seed=1), rows, 3 * H * DH, H * DH, H, DH, DHp, 3, PA(pl)],
             atol=1e-6)
    run_both("ctts_pad_heads_planes", [g(rows, H * DH), None, rows, H * DH, 0, H, DH, DHp, 2, PA(pl[:2])], atol=1e-4)


def test_pitch_frame_and_phoneme_kernels():
    B, S, M = 3, 9, 30
    n = B * M
    gen = torch.Generator().manual_seed(0)
    mel2ph = torch.randint(0, S + 1, (B, M), generator=gen)
    mel2ph, _ = torch.sort(mel2ph, dim=1)
    pred = g(B, M, 2) * 2 + torch.tensor([7.0, 0.0])
    run_both("ctts_frame_pitch", [pred, 2, None, None, mel2ph, 1, n, torch.zeros(n), torch.zeros(n), torch.zeros(n, dtype=torch.long)],
             atol=1e-3, rtol=1e-5)
    f0t = g(B, M) + 7.0
    run_both("ctts_frame_pitch", [None, 0, f0t, (g(B, M, seed=1) > 0).float(), mel2ph, 1, n, torch.zeros(n), torch.zeros(n),
                                  torch.zeros(n, dtype=torch.long)], atol=1e-3, rtol=1e-5)
    idx_ph = torch.randint(1, 255, (B, S), generator=gen)
    run_both("ctts_gather_index", [idx_ph, mel2ph, B, S, M, torch.zeros(B, M, dtype=torch.long)])
    m2 = mel2ph.clamp(min=1)
    run_both("ctts_phoneme_pitch", [g(B, M) + 7, m2, torch.tensor([9, 7, 9]), torch.tensor([30, 21, 30]), B, S, M, torch.zeros(B, S)],
             atol=1e-5)
