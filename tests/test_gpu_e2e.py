"""GPU: the whole forward path (model.forward -> C ABI -> sm_100a kernels) against the reference's golden
fixtures and against the CPU oracle.

north_star tolerance: mel outputs within 1e-3 abs + 1e-2 rel (fp32); LengthRegulator indices bit-exact.
"""
import os
import sys

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden"))
import cases  # noqa: E402
import ctts_b200  # noqa: E402
from oracle import ctts_oracle as O  # noqa: E402

DEV = "cuda:0"
MEL_ATOL, MEL_RTOL = 1e-3, 1e-2
EXACT = ("d_rounded", "mel_lens", "src_lens", "src_masks", "mel_masks", "p_targets.mel2ph",
         "attn_outs.1", "attn_outs.2")  # attn_hard (MAS path) and attn_hard_dur are integer-valued: bit-exact


def to_dev(v):
    if torch.is_tensor(v):
        return v.to(DEV)
    if isinstance(v, dict):
        return {k: to_dev(x) for k, x in v.items()}
    return v


def run_case(name):
    (p, m, t), sd, batch = cases.build_case(name)
    net = ctts_b200.CompTransTTS(p, m, t).eval()
    net.load_state_dict(sd, strict=True)
    net.to(DEV)
    args, kw = cases.call_kwargs(batch)
    out = net(*[to_dev(a) for a in args], **{k: to_dev(v) for k, v in kw.items()})
    torch.cuda.synchronize()
    return out, (p, m, t), sd, batch


def check_against(flat, gold, prefix):
    n = 0
    for key in gold:
        if not key.startswith(prefix):
            continue
        k = key[len(prefix):]
        assert k in flat, "output lacks %s" % k
        a, b = np.asarray(gold[key]), flat[k]
        assert a.shape == b.shape, (k, a.shape, b.shape)
        if k in EXACT or a.dtype.kind in "biu":
            assert np.array_equal(a, b), "%s must be bit-exact" % k
        elif k in ("mel", "postnet_mel"):
            np.testing.assert_allclose(b, a, atol=MEL_ATOL, rtol=MEL_RTOL, err_msg=k)
        else:
            np.testing.assert_allclose(b, a, atol=1e-3, rtol=1e-3, err_msg=k)
        n += 1
    return n


@pytest.mark.parametrize("name", sorted(cases.CASES))
def test_forward_matches_reference_golden(name, golden_dir):
    gold = np.load(os.path.join(golden_dir, name + ".npz"))
    out, _, _, _ = run_case(name)
    flat = cases.flatten_outputs(out)
    assert check_against(flat, {k: gold[k] for k in gold.files}, "ref.") >= 10
    # the fp32 path should in fact be far inside the tolerance
    scale = max(1.0, float(np.abs(gold["ref.postnet_mel"]).max()) / 4.0)
    assert np.abs(flat["postnet_mel"] - gold["ref.postnet_mel"]).max() < 2e-4 * scale


def test_forward_matches_oracle_batch16():
    """BASELINE configs[1] shape (B 16, S 100..70) at 4 frames / phoneme against the CPU oracle."""
    p, m, t = ctts_b200.builtin_configs("LJSpeech", learn_alignment=False)
    from ctts_b200 import spec, synth
    sd = synth.synthetic_state_dict(spec.parameter_spec(p, m)[0], pin_frames_per_phoneme=4)
    batch = synth.ljspeech_batch(batch=16, s_max=100, s_step=2, mode="infer")
    net = ctts_b200.CompTransTTS(p, m, t).eval()
    net.load_state_dict(sd)
    net.to(DEV)
    args, _ = cases.call_kwargs(batch)
    out = net(*[to_dev(a) for a in args])
    with torch.no_grad():
        ref = O.comp_trans_tts_forward(sd, p, m, t, *args)
    assert torch.equal(out[9].cpu(), ref[9]) and torch.equal(out[5].cpu(), ref[5])
    assert out[0].shape == (16, 400, 80)
    pidx, pref = O.f0_to_coarse(out[2]["f0_denorm"].cpu()), O.f0_to_coarse(ref[2]["f0_denorm"])
    assert int((pidx != pref).sum()) == 0, "pitch bucket flips"
    for i in (0, 1):
        np.testing.assert_allclose(out[i].cpu().numpy(), ref[i].numpy(), atol=MEL_ATOL, rtol=MEL_RTOL)


def test_longest_utterance_is_independent_and_controls_work():
    """Utterance 0 (the longest: no padding anywhere) must not depend on what else is in the batch, bit for bit;
    d_control scales the regulated length.  (Shorter utterances DO depend on the padded length in the reference --
    CWT normalisation and PostNet run over padded frames, SURVEY.md H4 -- and the golden tests cover that.)"""
    (p, m, t), sd, batch = cases.build_case("fs2_infer_c1")
    net = ctts_b200.CompTransTTS(p, m, t).eval()
    net.load_state_dict(sd)
    net.to(DEV)
    both = net(batch["speakers"].to(DEV), batch["texts"].to(DEV), batch["src_lens"].to(DEV), batch["max_src_len"])
    other = batch["texts"].clone()
    other[1, :50] = other[1, :50].flip(0)
    swapped = net(batch["speakers"].to(DEV), other.to(DEV), batch["src_lens"].to(DEV), batch["max_src_len"])
    assert torch.equal(both[1][0], swapped[1][0])
    assert not torch.equal(both[1][1], swapped[1][1])
    slow = net(batch["speakers"].to(DEV), batch["texts"].to(DEV), batch["src_lens"].to(DEV), batch["max_src_len"],
               d_control=2.0)
    assert slow[9].tolist() == [2 * v for v in both[9].tolist()]


def test_bench_workload_matches_oracle():
    """The EXACT benchmark workload (bench.build_workload(seed=0): B 16, S 100..70, 8 frames / phoneme, M 800, 10 880 valid
    frames) against the CPU oracle: lengths and durations bit-exact, zero pitch-bucket flips, mels within north_star."""
    import bench
    (p, m, t), sd, batch, frames = bench.build_workload(seed=0)
    assert frames == 10880
    net = ctts_b200.CompTransTTS(p, m, t).eval()
    net.load_state_dict(sd)
    net.to(DEV)
    args, _ = cases.call_kwargs(batch)
    outs = [net(*[to_dev(a) for a in args]) for _ in range(3)]       # eager, captured, replayed
    with torch.no_grad():
        ref = O.comp_trans_tts_forward(sd, p, m, t, *args)
    for out in outs:
        assert out[0].shape == (16, 800, 80)
        assert torch.equal(out[9].cpu(), ref[9]) and torch.equal(out[5].cpu(), ref[5])
        pidx, pref = O.f0_to_coarse(out[2]["f0_denorm"].cpu()), O.f0_to_coarse(ref[2]["f0_denorm"])
        assert int((pidx != pref).sum()) == 0, "pitch bucket flips"
        for i in (0, 1):
            np.testing.assert_allclose(out[i].cpu().numpy(), ref[i].numpy(), atol=MEL_ATOL, rtol=MEL_RTOL)
    assert torch.equal(outs[1][1], outs[2][1])


@pytest.mark.parametrize("name", ["fs2_infer_c1", "fs2_teacher", "fs2_unsup", "transformer_teacher", "fastformer_infer",
                                  "conformer_infer", "fastformer_vctk_unsup", "fs2_liu2021_infer"])
def test_cuda_graph_replay_equals_eager(name, monkeypatch):
    """The default production path captures the forward on the second call of a shape and replays it from the third on.
    Replayed outputs must equal an eager (CTTS_CUDA_GRAPHS=0) module bit for bit -- also when the input VALUES change at
    the same shapes -- and outputs handed out earlier must not be overwritten by later replays."""
    (p, m, t), sd, batch = cases.build_case(name)

    def make(graphs):
        monkeypatch.setenv("CTTS_CUDA_GRAPHS", "1" if graphs else "0")
        net = ctts_b200.CompTransTTS(p, m, t).eval()
        net.load_state_dict(sd, strict=True)
        return net.to(DEV)

    def variant(i):
        """the same shapes, different token values (valid region only; the padding stays 0)"""
        b = {k: (v.clone() if torch.is_tensor(v) else v) for k, v in batch.items()}
        if i:
            tx = b["texts"].clone()
            nz = tx != 0
            tx[nz] = (tx[nz] + 7 * i - 1) % 360 + 1
            b["texts"] = tx
        return b

    def call(net, b):
        args, kw = cases.call_kwargs(b)
        out = net(*[to_dev(a) for a in args], **{k: to_dev(v) for k, v in kw.items()})
        torch.cuda.synchronize()
        return out

    free_running = "d_targets" not in batch and "attn_priors" not in batch
    g_net, e_net = make(True), make(False)
    kept = []
    for i in (0, 0, 0, 1, 2, 0):
        if i and free_running:
            continue        # different tokens would change the regulated length, i.e. the shape key of stage B
        go, eo = call(g_net, variant(i)), call(e_net, variant(i))
        fg, fe = cases.flatten_outputs(go), cases.flatten_outputs(eo)
        assert sorted(fg) == sorted(fe)
        for k in fg:
            assert np.array_equal(fg[k], fe[k]), "%s differs between graph replay and eager (call with variant %d)" % (k, i)
        kept.append((go[1], go[1].clone()))
    for live, snapshot in kept:
        assert torch.equal(live, snapshot), "an earlier output was overwritten by a later graph replay"
    assert len(g_net._graphs.entries) >= 1


def test_positional_table_regrowth_keeps_captured_graphs_valid(monkeypatch):
    """A longer utterance regrows the cached sinusoid table; graphs captured before still read the old one."""
    monkeypatch.setenv("CTTS_CUDA_GRAPHS", "1")
    p, m, t = ctts_b200.builtin_configs("LJSpeech", learn_alignment=False)
    m["transformer_fs2"]["encoder_layer"] = 1
    m["transformer_fs2"]["decoder_layer"] = 1
    from ctts_b200 import spec, synth

    def run(net, frames_pp, n_calls):
        batch = synth.ljspeech_batch(batch=2, s_max=100, s_step=10, mode="infer", seed=3)
        args, _ = cases.call_kwargs(batch)
        return [net(*[to_dev(a) for a in args], d_control=float(frames_pp) / 8.0)[1].clone() for _ in range(n_calls)]

    sd = synth.synthetic_state_dict(spec.parameter_spec(p, m)[0], pin_frames_per_phoneme=8)
    net = ctts_b200.CompTransTTS(p, m, t).eval()
    net.load_state_dict(sd)
    net.to(DEV)
    first = run(net, 24, 3)                  # M = 2400 > 2048 rows: table sized for it, graphs captured
    n_tables = len(net._prepared.pe_retired)
    run(net, 48, 2)                          # M = 4800: the table is regrown
    assert len(net._prepared.pe_retired) > n_tables
    again = run(net, 24, 2)                  # replays the graphs captured against the OLD table
    monkeypatch.setenv("CTTS_CUDA_GRAPHS", "0")
    eager = ctts_b200.CompTransTTS(p, m, t).eval()
    eager.load_state_dict(sd)
    eager.to(DEV)
    want = run(eager, 24, 1)[0]
    for got in first + again:
        assert torch.equal(got, want)


def test_training_mode_needs_targets():
    """model.train() runs the training step (tests/test_gpu_train.py); without targets there is nothing to train on."""
    (p, m, t), sd, batch = cases.build_case("fs2_infer_c1")
    net = ctts_b200.CompTransTTS(p, m, t).train().to(DEV)
    with pytest.raises(AssertionError):
        net(batch["speakers"].to(DEV), batch["texts"].to(DEV), batch["src_lens"].to(DEV), batch["max_src_len"])
