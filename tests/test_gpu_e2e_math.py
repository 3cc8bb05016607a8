"""GPU: both arithmetic modes of the decoder (tcgen05 bf16x3 and CUDA-core fp32) against the reference fixtures."""
import os
import sys

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden"))
import cases  # noqa: E402
import ctts_b200  # noqa: E402

DEV = "cuda:0"


@pytest.mark.parametrize("math", ["bf16x3", "fp32"])
@pytest.mark.parametrize("name", sorted(cases.CASES))
def test_decoder_math_modes(name, math, golden_dir, monkeypatch):
    monkeypatch.setenv("CTTS_DECODER_MATH", math)
    monkeypatch.setenv("CTTS_ENCODER_MATH", "bf16x6" if math == "bf16x3" else "fp32")
    gold = np.load(os.path.join(golden_dir, name + ".npz"))
    (p, m, t), sd, batch = cases.build_case(name)
    net = ctts_b200.CompTransTTS(p, m, t).eval()
    assert net.decoder_math == math
    net.load_state_dict(sd, strict=True)
    net.to(DEV)
    args, kw = cases.call_kwargs(batch)
    mv = lambda v: v.to(DEV) if torch.is_tensor(v) else ({k: x.to(DEV) for k, x in v.items()} if isinstance(v, dict) else v)
    out = net(*[mv(a) for a in args], **{k: mv(v) for k, v in kw.items()})
    for i, key in ((0, "ref.mel"), (1, "ref.postnet_mel")):
        got = out[i].cpu().numpy()
        np.testing.assert_allclose(got, gold[key], atol=1e-3, rtol=1e-2, err_msg="%s (%s)" % (key, math))
        # both modes are in fact far inside the north_star tolerance
        scale = max(1.0, float(np.abs(gold[key]).max()) / 4.0)
        assert np.abs(got - gold[key]).max() < (3e-4 if math == "bf16x3" else 2e-4) * scale
    np.testing.assert_allclose(out[0].cpu().numpy()[:, ::cases.TAP_STRIDE][:, :, :0], gold["ref.mel"][:, ::cases.TAP_STRIDE][:, :, :0])
