"""GPU: every entry point of the training step (include/ctts_b200.h, "TRAINING STEP") against its CPU restatement
(oracle/capi_emulator.py) on the same seeded inputs.  The real kernel mutates its output arguments on the device, the
restatement mutates CPU copies; afterwards every tensor argument must agree."""
import math
import os
import sys

import numpy as np
import pytest
import torch

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from ctts_b200 import capi
from oracle import capi_emulator as emu

pytestmark = pytest.mark.gpu
DEV = torch.device("cuda:0")


from gpu_harness import II, LL, PA, Scratch, g, planes, run_both  # noqa: E402


# ---------------------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("shape", [(1, 1, 70, 50, 33), (3, 1, 130, 65, 48), (6, 2, 40, 24, 128)])
def test_gemm_generic_plain_and_batched(shape):
    Z, zmod, M, N, K = shape
    a, b = g(Z, M, K), g(Z, N, K, seed=1)
    y = g(Z, M, N, seed=2)
    zo, zi = Z // zmod, zmod
    for acc in (0, 1):
        run_both("ctts_gemm_generic", [a, b, y, Z, zmod, M, N, K, LL(zi * M * K, M * K, K, 1, 0), LL(zi * N * K, N * K, K, 1, 0),
                                       LL(zi * M * N, M * N, N, 1), 0, 0, 0, 0.7, acc], atol=2e-4)


def test_gemm_generic_transposed_operands():
    M, N, K = 37, 29, 53
    a, b, y = g(K, M), g(K, N, seed=1), torch.zeros(N, M)          # A[m,k] = a[k,m]; y stored transposed
    run_both("ctts_gemm_generic", [a, b, y, 1, 1, M, N, K, LL(0, 0, 1, M, 0), LL(0, 0, 1, N, 0), LL(0, 0, 1, M), 0, 0, 0, 1.0, 0],
             atol=2e-4)


@pytest.mark.parametrize("taps", [3, 9])
def test_gemm_generic_conv_weight_gradient(taps):
    """reduction over (utterance, time) with a tap shift: dw[n, tap*Cin + c] = sum dz[b,t,n] x[b,t+tap-pad,c]"""
    B, T, N, Cin = 3, 21, 20, 16
    dz, x = g(B, T, N), g(B, T, Cin, seed=1)
    out = torch.zeros(N, taps * Cin)
    run_both("ctts_gemm_generic", [dz, x, out, taps, 1, N, Cin, B * T, LL(0, 0, 1, N, T * N), LL(0, 0, 1, Cin, T * Cin),
                                   LL(Cin, 0, taps * Cin, 1), T, -(taps // 2), 1, 1.0, 0], atol=2e-4)
    w = out  # check against autograd through conv1d on the CPU as well
    xx = x.clone()
    wt = torch.zeros(N, Cin, taps, requires_grad=True)
    torch.nn.functional.conv1d(xx.transpose(1, 2), wt, padding=taps // 2).transpose(1, 2).backward(dz)
    ref = wt.grad.permute(0, 2, 1).reshape(N, taps * Cin)
    cpu = torch.zeros(N, taps * Cin)
    emu.ctts_gemm_generic(dz, x, cpu, taps, 1, N, Cin, B * T, [0, 0, 1, N, T * N], [0, 0, 1, Cin, T * Cin],
                          [Cin, 0, taps * Cin, 1], T, -(taps // 2), 1, 1.0, 0, 0)
    assert torch.allclose(cpu, ref, atol=1e-4)


@pytest.mark.parametrize("act", [0, 1, 2, 3, 4])
def test_act_bwd(act):
    B, T, N = 3, 50, 70
    dy, ref = g(B, T, N), g(B, T, N, seed=1)
    if act == 3:
        ref = torch.tanh(ref)
    lens = torch.tensor([50, 31, 7])
    dz, db = torch.zeros(B, T, N), g(N, seed=3)
    run_both("ctts_act_bwd", [dy, ref if act else None, act, 0.33, lens, 1, T, B * T, N, dz, db], atol=2e-5)
    run_both("ctts_act_bwd", [dy, ref if act else None, act, 1.0, None, 1, T, B * T, N, dy.clone(), None])


def test_act_bwd_row_broadcast_form():
    B, T, C = 4, 37, 48
    dy, out = g(B, T, C), torch.zeros(B, C)
    run_both("ctts_act_bwd", [dy, None, 0, 1.0, None, B, T, T, C, None, out], atol=2e-5)


@pytest.mark.parametrize("C,lens", [(256, True), (128, False), (80, True)])
def test_layernorm_bwd(C, lens):
    B, T = 3, 41
    x, gamma, dy = g(B, T, C), 1 + 0.1 * g(C, seed=1), g(B, T, C, seed=2)
    ln = torch.tensor([41, 20, 3]) if lens else None
    for acc in (0, 1):
        run_both("ctts_layernorm_bwd", [x, gamma, dy, 1e-5, ln, B, T, C, g(B, T, C, seed=4), acc, g(C, seed=5), g(C, seed=6)],
                 atol=5e-5, rtol=1e-3)


def test_mask_axpy_rowscale_copy_rows():
    B, T, C = 3, 19, 32
    run_both("ctts_mask_rows", [g(B, T, C), torch.tensor([19, 4, 0]), B, T, C])
    run_both("ctts_mask_rows", [g(B, T, 7), torch.tensor([19, 4, 0]), B, T, 7])
    run_both("ctts_axpy", [g(1000), 0.5, 1000, 1, g(1000, seed=1)])
    run_both("ctts_axpy", [g(1001), 2.0, 1001, 0, g(1001, seed=1)])
    run_both("ctts_rowscale_axpy", [g(57, 24), g(57, seed=1), -0.3, 57, 24, 1, g(57, 24, seed=2)])
    run_both("ctts_copy_rows", [g(B, T, C), T * C, B, C, torch.zeros(B, C), C, 0])
    run_both("ctts_copy_rows", [g(B, C), C, B, C, g(B, T, C, seed=1), T * C, 1])


def test_scatter_add_rows_and_length_expand_bwd():
    B, T, C, V = 3, 40, 64, 17
    gen = torch.Generator().manual_seed(0)
    idx = torch.randint(0, V, (B, T), generator=gen)
    run_both("ctts_scatter_add_rows", [g(B, T, C), idx, torch.tensor([40, 22, 5]), T, B * T, C, V, 0, 16.0, torch.zeros(V, C)],
             atol=1e-4)
    run_both("ctts_scatter_add_rows", [g(B, T, C), idx, None, T, B * T, C, V, -1, 1.0, g(V, C, seed=1)], atol=1e-4)
    S, M = 9, 30
    dur = torch.tensor([[3, 0, 5, 1, 1, 7, 2, 0, 4], [1, 1, 1, 1, 1, 1, 1, 1, 1], [0, 0, 0, 0, 40, 0, 0, 0, 0]])
    cum = torch.cumsum(dur, 1).int()
    for acc in (0, 1):
        run_both("ctts_length_expand_bwd", [g(B, M, C), cum, B, S, C, M, acc, g(B, S, C, seed=1)], atol=1e-5)


@pytest.mark.parametrize("pos_mode", [0, 1])
def test_add_positions_bwd(pos_mode):
    B, T, C = 3, 70, 64
    x = g(B, T, C)
    x[1, 40:, 0] = 0
    pe = g(T + 2, C, seed=1)
    run_both("ctts_add_positions_bwd", [g(B, T, C, seed=2), x, pe, T + 2, torch.tensor([70, 50, 9]), B, T, C, pos_mode,
                                        torch.tensor([0.25])], atol=1e-3, rtol=1e-4)


def test_masked_softmax_and_backward():
    B, H, T = 2, 2, 45
    Z = B * H
    S = g(Z, T, T)
    lens = torch.tensor([45, 17])
    P = torch.zeros(Z, T, T)
    run_both("ctts_masked_softmax", [S, lens, H, Z, T, T, T, 1, P], atol=1e-6)
    emu.ctts_masked_softmax(S, lens, H, Z, T, T, T, 1, P, 0)
    run_both("ctts_softmax_bwd", [P, g(Z, T, T, seed=1), Z, T, T, T, 0.5, torch.zeros(Z, T, T)], atol=1e-6)
    run_both("ctts_masked_softmax", [S, None, 1, Z, T, T, T, 0, P], atol=1e-6)


@pytest.mark.parametrize("act", [0, 3, 4])
def test_batchnorm_training(act):
    rows, C = 333, 80
    x = g(rows, C) * 2 + 0.5
    mean, var = torch.zeros(C), torch.zeros(C)
    run_both("ctts_bn_stats", [x, rows, C, mean, var], atol=1e-5)
    emu.ctts_bn_stats(x, rows, C, mean, var, 0)
    gamma, beta = 1 + 0.1 * g(C, seed=1), 0.1 * g(C, seed=2)
    y = torch.zeros(rows, C)
    pl = [torch.zeros(rows, C, dtype=torch.bfloat16) for _ in range(2)]
    run_both("ctts_bn_act_fwd", [x, mean, var, gamma, beta, 1e-5, act, rows, C, y, 2, PA(pl)], atol=1e-5)
    run_both("ctts_bn_update_running", [mean, var, rows, 0.1, C, g(C, seed=3), g(C, seed=4).abs(), torch.tensor([5])])
    run_both("ctts_bn_bwd", [g(rows, C, seed=5), x, mean, var, gamma, beta, 1e-5, act, rows, C, torch.zeros(rows, C), g(C, seed=6),
                             g(C, seed=7), Scratch(torch.zeros(2 * C))], atol=1e-4, rtol=1e-3)


def test_dropout_is_the_philox_stream_and_keeps_1_minus_p():
    n = 100003
    x = torch.ones(n)
    y = torch.zeros(n)
    run_both("ctts_dropout", [x, n, 0.3, 1234, 7, None, y], atol=1e-6)
    run_both("ctts_dropout", [x, n, 0.3, 1234, 2, torch.tensor([5]), y], atol=1e-6)
    yg = torch.zeros(n, device=DEV)
    capi.call("ctts_dropout", x.to(DEV), n, 0.3, 1234, 7, None, yg, torch.cuda.current_stream().cuda_stream)
    kept = (yg > 0).float().mean().item()
    assert abs(kept - 0.7) < 0.01
    assert abs(yg.mean().item() - 1.0) < 0.02          # scaled by 1 / (1 - p)
    y2 = torch.zeros(n, device=DEV)
    capi.call("ctts_dropout", x.to(DEV), n, 0.3, 1234, 8, None, y2, torch.cuda.current_stream().cuda_stream)
    assert (yg != y2).float().mean().item() > 0.3      # a different offset draws a different mask


def test_weight_relayouts():
    N, Cin, taps = 20, 12, 5
    w = g(N, Cin, taps)
    run_both("ctts_pack_conv_weight_dgrad", [w, N, Cin, taps, torch.zeros(Cin, taps * N)])
    run_both("ctts_unpack_conv_wgrad", [g(N, taps * Cin), N, Cin, taps, 1, g(N, Cin, taps, seed=1)])
    run_both("ctts_act_fwd", [g(1000), 1000, 2, torch.zeros(1000), 2, PA([torch.zeros(1000, dtype=torch.bfloat16)] * 2)],
             atol=1e-5)
    x = g(777)
    run_both("ctts_merge_planes", [3, PA(planes(x, 3)), 777, torch.zeros(777)], atol=1e-7)


def test_split_transpose():
    Z, R, C = 3, 45, 70
    Rp = 48
    x = g(Z, R, C)
    pl = [torch.zeros(Z, C, Rp, dtype=torch.bfloat16) for _ in range(2)]
    run_both("ctts_split_transpose", [x, Z, R, C, C, 0, Rp, 1, 2, PA(pl)], atol=0, rtol=0)
    pl = [torch.zeros(Z, 5, C, Rp, dtype=torch.bfloat16) for _ in range(2)]
    run_both("ctts_split_transpose", [x, Z, R, C, C, 0, Rp, 5, 2, PA(pl)], atol=0, rtol=0)


@pytest.mark.parametrize("shape", [(2, 40, 64, 128, 1), (3, 100, 256, 1024, 9), (2, 33, 80, 160, 3), (16, 64, 256, 256, 1),
                                   (2, 50, 512, 80, 5), (16, 100, 256, 1, 1), (4, 50, 256, 11, 1), (16, 800, 256, 256, 1)])
def test_gemm_wgrad_tensor_core(shape):
    """tcgen05 weight gradient against the fp64 definition (and the emulator)."""
    B, T, Cin, N, taps = shape
    Tp = (T + 7) // 8 * 8
    dz, x = g(B, T, N), g(B, T, Cin, seed=1)
    dzT = [torch.zeros(B, N, Tp, dtype=torch.bfloat16) for _ in range(2)]
    xT = [torch.zeros(B, taps, Cin, Tp, dtype=torch.bfloat16) for _ in range(2)]
    emu.ctts_split_transpose(dz, B, T, N, N, 0, Tp, 1, 2, capi.ptr_array(dzT), 0)
    emu.ctts_split_transpose(x, B, T, Cin, Cin, 0, Tp, taps, 2, capi.ptr_array(xT), 0)
    out = g(N, taps * Cin, seed=2)
    scale = math.sqrt(B * T)
    for acc in (0, 1):
        run_both("ctts_gemm_wgrad", [2, PA(dzT), PA(xT), B, T, Tp, Cin, N, taps, 1.0, acc, out], atol=2e-4 * scale, rtol=1e-4)


@pytest.mark.parametrize("shape", [(2, 40, 128, 128, 1), (3, 100, 256, 1024, 9), (2, 33, 128, 160, 3), (16, 64, 256, 256, 1),
                                   (2, 50, 512, 80, 5), (4, 50, 256, 8, 1), (16, 800, 256, 256, 1), (2, 7, 1024, 256, 1)])
def test_gemm_wgrad_rowmajor(shape):
    """The weight gradient straight from the row-major planes (both operands MN-major, the tap a row offset of the x box)
    against the fp64 definition restated by the emulator; both accumulate modes."""
    B, T, Cin, N, taps = shape
    dz, x = g(B, T, N), g(B, T, Cin, seed=1)
    dzp = [torch.zeros(B, T, N, dtype=torch.bfloat16) for _ in range(2)]
    xp = [torch.zeros(B, T, Cin, dtype=torch.bfloat16) for _ in range(2)]
    emu.ctts_split_planes(dz, dz.numel(), 2, capi.ptr_array(dzp), 0)
    emu.ctts_split_planes(x, x.numel(), 2, capi.ptr_array(xp), 0)
    out = g(N, taps * Cin, seed=2)
    scale = math.sqrt(B * T)
    for acc in (0, 1):
        run_both("ctts_gemm_wgrad_rowmajor", [2, PA(dzp), PA(xp), B, T, Cin, N, taps, 1.0, acc, out], atol=2e-4 * scale, rtol=1e-4)


def test_gemm_wgrad_bench_shape_accumulation_error():
    """The FFN conv of the benchmark: K = 16 x 800 rows through ONE TMEM accumulator.  Measures the error against fp64."""
    B, T, Cin, N, taps = 16, 800, 256, 1024, 9
    Tp = T
    gen = torch.Generator(device="cuda").manual_seed(0)
    dz = torch.randn(B, T, N, device=DEV, generator=gen)
    x = torch.randn(B, T, Cin, device=DEV, generator=gen)
    st = torch.cuda.current_stream().cuda_stream
    dzT = [torch.empty(B, N, Tp, dtype=torch.bfloat16, device=DEV) for _ in range(2)]
    xT = [torch.empty(B, taps, Cin, Tp, dtype=torch.bfloat16, device=DEV) for _ in range(2)]
    capi.call("ctts_split_transpose", dz, B, T, N, N, 0, Tp, 1, 2, capi.ptr_array(dzT), st)
    capi.call("ctts_split_transpose", x, B, T, Cin, Cin, 0, Tp, taps, 2, capi.ptr_array(xT), st)
    out = torch.empty(N, taps * Cin, device=DEV)
    capi.call("ctts_gemm_wgrad", 2, capi.ptr_array(dzT), capi.ptr_array(xT), B, T, Tp, Cin, N, taps, 1.0, 0, out, st)
    torch.cuda.synchronize()
    w = torch.zeros(N, Cin, taps, device=DEV, dtype=torch.float64, requires_grad=True)
    torch.nn.functional.conv1d(x.double().transpose(1, 2), w, padding=taps // 2).transpose(1, 2).backward(dz.double())
    ref = w.grad.permute(0, 2, 1).reshape(N, taps * Cin)
    err = (out.double() - ref).abs().max().item()
    rms = ref.pow(2).mean().sqrt().item()
    print("wgrad bench shape: max err %.3g, rms of the gradient %.3g, ratio %.3g" % (err, rms, err / rms))
    assert err <= 2e-3 * rms


def test_gemm_batched_planes_matches_definition():
    Z, T, K, N = 4, 70, 64, 48
    a, w = g(Z, T, K), g(Z, N, K, seed=1)
    y = torch.zeros(Z, T, N)
    run_both("ctts_gemm_batched_planes", [2, PA(planes(a)), LL(K, T, Z, K, T * K), PA(planes(w)), LL(K, N, Z, K, N * K),
                                          II(1, 1, 0, 0, 1, 0, 0, 1, N), T * N, 0, 0.5, None, None, Z, T, K, N, y, None],
             atol=1e-4, rtol=1e-4)


def test_aligner_attention_bwd():
    B, M, S = 2, 37, 21
    lens = torch.tensor([21, 13])
    q, k = g(B, M, 80), g(B, S, 80, seed=1)
    prior = torch.rand(B, S, M, generator=torch.Generator().manual_seed(0))
    soft, logprob = torch.zeros(B, M, S), torch.zeros(B, M, S)
    emu.ctts_aligner_attention(q, k, prior, lens, 0.0005, B, M, S, 80, soft, logprob, 0)
    run_both("ctts_aligner_attention_bwd", [soft, logprob, prior, g(B, M, S, seed=2), g(B, M, S, seed=3), lens, B, M, S,
                                            torch.zeros(B, M, S)], atol=2e-5)
    run_both("ctts_aligner_attention_bwd", [soft, logprob, prior, None, g(B, M, S, seed=3), lens, B, M, S, torch.zeros(B, M, S)],
             atol=2e-5)


def test_block_specific_backward_kernels():
    rows, C = 123, 40
    run_both("ctts_glu_bwd", [g(rows, 2 * C), g(rows, C, seed=1), rows, C, torch.zeros(rows, 2 * C)])
    B, T, C, K = 2, 50, 48, 31
    x, w = g(B, T, C), g(C, K, seed=1) * 0.2
    run_both("ctts_dwconv", [x, w, K, B, T, C, torch.zeros(B, T, C)], atol=1e-5)
    run_both("ctts_dwconv_bwd", [g(B, T, C, seed=2), x, w, K, B, T, C, torch.zeros(B, T, C), g(C, K, seed=3)], atol=1e-4)
    Z, T = 3, 29
    run_both("ctts_relshift_bwd", [g(Z, T, T + 3), Z, T, T + 3, T, 16.0, torch.zeros(Z, T, T), torch.zeros(Z, T, T)])
    run_both("ctts_relshift_bwd", [g(Z, T, T + 3), Z, T, T + 3, 32, 16.0, torch.ones(Z, T, 32), torch.ones(Z, T, 32)])
    B, T, heads, hs = 2, 40, 128, 2
    lens = torch.tensor([40, 25])
    run_both("ctts_fastformer_pool_bwd", [g(B, T, heads), g(B, T, heads * hs, seed=1), lens, g(B, heads * hs, seed=2), B, T, heads,
                                          hs, torch.zeros(B, T, heads), torch.zeros(B, T, heads * hs)], atol=1e-5)
    run_both("ctts_mul_bwd", [g(B, T, 32), g(B, T, 32, seed=1), g(B, 32, seed=2), 1, lens, B, T, 32, torch.zeros(B, T, 32),
                              torch.zeros(B, 32)], atol=1e-5)
    run_both("ctts_mul_bwd", [g(B, T, 32), g(B, T, 32, seed=1), g(B, T, 32, seed=2), 0, None, B, T, 32, torch.zeros(B, T, 32),
                              torch.zeros(B, T, 32)], atol=1e-5)


@pytest.mark.parametrize("reverse", [0, 1])
def test_gru_bwd(reverse):
    B, T, H = 2, 17, 32
    gi = g(B, T, 3 * H)
    whh, bhh = g(3 * H, H, seed=1) * 0.2, g(3 * H, seed=2) * 0.1
    out, hf = torch.zeros(B, T, 2 * H), torch.zeros(B, 2 * H)
    emu.ctts_gru_bidir(gi, gi, whh, bhh, whh, bhh, B, T, H, out, hf, 0)
    run_both("ctts_gru_bwd", [gi, whh, bhh, out, 2 * H, reverse * H, g(B, T, 2 * H, seed=3), g(B, 2 * H, seed=4).view(-1)[reverse * H:],
                              2 * H, B, T, H, reverse, torch.zeros(B, T, 3 * H), torch.zeros(B, T, 3 * H)], atol=2e-5, rtol=1e-3)


def test_reference_encoder_helper_kernels():
    """AddCoords / im2col / col2im of the 3x3 stride-(1,2) convolutions / last-two-dims permute (liu2021 training)."""
    N, H, W, C = 2, 13, 10, 8
    run_both("ctts_add_coords", [g(N, H, W), N, H, W, torch.zeros(N, H, W, 4)], atol=1e-6)
    Wo = (W + 2 - 3) // 2 + 1
    run_both("ctts_im2col_3x3_s12", [g(N, H, W, C), N, H, W, C, torch.zeros(N * H * Wo, 9 * C)], atol=0, rtol=0)
    run_both("ctts_col2im_3x3_s12", [g(N * H * Wo, 9 * C), N, H, W, C, torch.zeros(N, H, W, C)], atol=1e-5)
    W = 5      # odd width: the last output column reads one real and one padded input column
    Wo = (W + 2 - 3) // 2 + 1
    run_both("ctts_im2col_3x3_s12", [g(N, H, W, C), N, H, W, C, torch.zeros(N * H * Wo, 9 * C)], atol=0, rtol=0)
    run_both("ctts_col2im_3x3_s12", [g(N * H * Wo, 9 * C), N, H, W, C, torch.zeros(N, H, W, C)], atol=1e-5)
    run_both("ctts_permute_last2", [g(7, 3, 128), 7, 3, 128, torch.zeros(7, 128, 3)], atol=0, rtol=0)


@pytest.mark.parametrize("act,N", [(0, 256), (2, 1024), (1, 80), (3, 12)])
def test_act_bwd_planes_fused(act, N):
    B, T = 3, 45
    Tp = 48
    dy, ref = g(B, T, N), g(B, T, N, seed=1)
    if act == 3:
        ref = torch.tanh(ref)
    lens = torch.tensor([45, 20, 3])
    zp = [torch.zeros(B, T, N, dtype=torch.bfloat16) for _ in range(2)]
    zt = [torch.zeros(B, N, Tp, dtype=torch.bfloat16) for _ in range(2)]
    run_both("ctts_act_bwd_planes", [dy, ref if act else None, act, 0.5, lens, B, T, N, Tp, torch.zeros(B, T, N), 2, PA(zp), PA(zt),
                                     g(N, seed=3)], atol=3e-5)
    run_both("ctts_act_bwd_planes", [dy, ref if act else None, act, 1.0, None, B, T, N, Tp, None, 2, PA(zp), PA(zt), None], atol=3e-5)


def test_dropout_add_fused():
    B, T, C = 3, 20, 64
    lens = torch.tensor([20, 11, 0])
    run_both("ctts_dropout_add", [g(B, T, C), g(B, T, C, seed=1), lens, B, T, C, 0.1, 77, 3, torch.tensor([9]), torch.zeros(B, T, C)],
             atol=1e-6)
    run_both("ctts_dropout_add", [g(B, T, C), g(B, T, C, seed=1), None, B, T, C, 0.5, 77, 3, None, torch.zeros(B, T, C)], atol=1e-6)
