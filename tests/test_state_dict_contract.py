"""CPU: constructor / state_dict contract of the drop-in module."""
import os

import pytest
import torch

import ctts_b200
from ctts_b200 import spec, synth


def test_fs2_state_dict_names_and_shapes():
    p, m, t = ctts_b200.builtin_configs("LJSpeech", learn_alignment=False)
    net = ctts_b200.CompTransTTS(p, m, t)
    sd = net.state_dict()
    assert sd["encoder.layers.0.op.self_attn.in_proj_weight"].shape == (768, 256)
    assert sd["encoder.layers.3.op.ffn.ffn_1.weight"].shape == (1024, 256, 9)
    assert sd["decoder.layers.5.op.ffn.ffn_2.weight"].shape == (256, 1024)
    assert sd["variance_adaptor.energy_bins"].shape == (255,)
    assert sd["variance_adaptor.cwt_predictor.1.linear.weight"].shape == (11, 256)
    assert sd["postnet.convolutions.4.1.num_batches_tracked"].dtype == torch.long
    assert "variance_adaptor.aligner.key_proj.0.conv.weight" not in sd
    assert sum(v.numel() for v in net.parameters()) == 35094225  # probed on the reference (learn_alignment False)
    assert not net.get_parameter("variance_adaptor.energy_bins").requires_grad


def test_aligner_parameters_follow_learn_alignment():
    p, m, t = ctts_b200.builtin_configs("LJSpeech", learn_alignment=True)
    sd = ctts_b200.CompTransTTS(p, m, t).state_dict()
    assert sd["variance_adaptor.aligner.key_proj.0.conv.weight"].shape == (512, 256, 3)
    assert sd["variance_adaptor.aligner.query_proj.4.conv.weight"].shape == (80, 80, 1)


def test_synthetic_state_dict_loads_strict_and_is_deterministic():
    p, m, t = ctts_b200.builtin_configs("LJSpeech", learn_alignment=False)
    entries, _, _ = spec.parameter_spec(p, m)
    a = synth.synthetic_state_dict(entries)
    b = synth.synthetic_state_dict(entries)
    assert all(torch.equal(a[k], b[k]) for k in a)
    net = ctts_b200.CompTransTTS(p, m, t)
    net.load_state_dict(a, strict=True)


def test_unknown_block_type_raises_not_implemented():
    p, m, t = ctts_b200.builtin_configs("LJSpeech")
    m["block_type"] = "no_such_block"
    with pytest.raises(NotImplementedError):  # model/CompTransTTS.py:31-32
        ctts_b200.CompTransTTS(p, m, t)


@pytest.mark.skipif(not os.path.isdir("/root/reference/model"), reason="reference tree not mounted")
def test_reference_accepts_our_state_dict():
    import subprocess, sys, textwrap
    code = textwrap.dedent("""
        import sys
        sys.path.insert(0, %r); sys.path.insert(0, %r)
        from oracle.ref_import import import_reference, reference_configs
        import ctts_b200
        ref_model, _ = import_reference()
        for learn in (False, True):
            rp, rm, rt = reference_configs("LJSpeech")
            rm["duration_modeling"]["learn_alignment"] = learn
            ref = ref_model.CompTransTTS(rp, rm, rt)
            p, m, t = ctts_b200.builtin_configs("LJSpeech", learn_alignment=learn)
            ours = ctts_b200.CompTransTTS(p, m, t)
            ref.load_state_dict(ours.state_dict(), strict=True)
            ours.load_state_dict(ref.state_dict(), strict=True)
        print("OK")
    """) % (os.path.dirname(os.path.dirname(os.path.abspath(__file__))),
            os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "comprehensive-transformer-tts_b200"))
    out = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=600)
    assert "OK" in out.stdout, out.stderr[-2000:]
