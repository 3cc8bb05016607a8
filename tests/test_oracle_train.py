"""CPU: the training-mode oracle (forward with batch-statistics BatchNorm and the liu2021 reference encoders, backward
by torch.autograd through the functional restatement) against fixtures produced by the reference in model.train() mode
with every dropout probability 0 (tests/golden/make_golden_train.py).  This pins the oracle for the SURVEY.md section 8
rows that are still to be built on the GPU (training step, A19 reference encoders)."""
import os
import sys

import numpy as np
import pytest
import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden"))
import cases  # noqa: E402
from oracle import ctts_oracle as O  # noqa: E402


def run_oracle_train(name, frozen=()):
    (p, m, t), sd, batch = cases.build_case(name)
    P = {k: (v.clone().requires_grad_(True) if v.is_floating_point() and "running_" not in k and k not in frozen
             else v.clone()) for k, v in sd.items()}
    from ctts_b200 import spec
    for k, _, _, init in spec.parameter_spec(p, m)[0]:     # tied entries are ONE tensor under several names, as in the
        if init.startswith("tie:"):                        # reference, so its gradient accumulates over all uses
            P[k] = P[init[4:]]
    args, kw = cases.call_kwargs(batch, cases.TRAIN_CASES[name].get("step"))
    stats = {}
    out = O.comp_trans_tts_forward(P, p, m, t, *args, training=True, stats_out=stats, **kw)
    loss = cases.train_objective(out)
    loss.backward()
    return P, out, loss, stats


@pytest.mark.parametrize("name", sorted(cases.TRAIN_CASES))
def test_training_oracle_matches_reference(name, golden_dir):
    gold = np.load(os.path.join(golden_dir, name + ".npz"))
    frozen = {k[7:] for k in gold.files if k.startswith("frozen.")}
    P, out, loss, stats = run_oracle_train(name, frozen)
    flat = cases.flatten_outputs(out)
    n_out = 0
    for key in gold.files:
        if key.startswith("ref."):
            k = key[4:]
            assert k in flat, "oracle output lacks %s" % k
            a, b = gold[key], flat[k]
            assert a.shape == b.shape, (k, a.shape, b.shape)
            if a.dtype.kind in "biu":
                assert np.array_equal(a, b), k
            else:
                np.testing.assert_allclose(b, a, atol=3e-5, rtol=2e-4, err_msg=k)
            n_out += 1
    assert n_out >= 10
    assert abs(loss.item() - float(gold["loss"])) <= 1e-4 * max(1.0, abs(float(gold["loss"])))

    # gradients: strided samples of every parameter's gradient + its norm
    n_grad = n_nonzero = 0
    for key in gold.files:
        if not key.startswith("grad."):
            continue
        k = key[5:]
        g = P[k].grad
        gf = (g if g is not None else torch.zeros_like(P[k])).reshape(-1)
        mine = gf[torch.from_numpy(cases.grad_sample_index(gf.numel()))].numpy()
        ref = gold[key]
        norm = float(gold["gnorm." + k])
        # summation-order noise scales with the tensor's gradient norm, not with the individual entry
        tol = 2e-4 * max(norm / max(gf.numel(), 1) ** 0.5, 1e-6) + 1e-6
        np.testing.assert_allclose(mine, ref, atol=20 * tol, rtol=2e-3, err_msg="grad " + k)
        assert abs(float(gf.double().norm()) - norm) <= 1e-3 * norm + 1e-6, "gradient norm of " + k
        n_grad += 1
        n_nonzero += int(norm > 0)
    assert n_grad >= 150 and n_nonzero >= n_grad - len(frozen) - 6   # frozen tables, unused CoordConv parent tensors

    # BatchNorm buffers after the step
    n_buf = 0
    for key in gold.files:
        if key.startswith("buf."):
            k = key[4:]
            if k in stats:
                np.testing.assert_allclose(stats[k].numpy(), gold[key], atol=1e-5, rtol=1e-4, err_msg=k)
                n_buf += 1
    assert n_buf >= 15


def test_training_oracle_is_the_eval_oracle_apart_from_batchnorm():
    """Without liu2021 the only forward difference between the two modes is PostNet's BatchNorm statistics: everything
    up to the mel_linear output must be identical, the post-net output must differ."""
    (p, m, t), sd, batch = cases.build_case("fs2_train")
    args, kw = cases.call_kwargs(batch)
    with torch.no_grad():
        a = O.comp_trans_tts_forward(sd, p, m, t, *args, **kw)
        b = O.comp_trans_tts_forward(sd, p, m, t, *cases.call_kwargs(batch)[0], training=True, **cases.call_kwargs(batch)[1])
    assert torch.equal(a[0], b[0]) and torch.equal(a[4], b[4])
    assert (a[1] - b[1]).abs().max().item() > 1e-3


@pytest.mark.parametrize("name", sorted(cases.TRAIN_CASES))
def test_drop_in_module_freezes_what_the_reference_freezes(name, golden_dir):
    """requires_grad flags of the drop-in module's parameters == the reference's (sinusoid tables and bucket edges are
    nn.Parameter(requires_grad=False) there): an optimizer built over model.parameters() must see the same set."""
    import ctts_b200
    gold = np.load(os.path.join(golden_dir, name + ".npz"))
    frozen = {k[7:] for k in gold.files if k.startswith("frozen.")}
    (p, m, t), _, _ = cases.build_case(name)
    net = ctts_b200.CompTransTTS(p, m, t)
    mine = {k for k, prm in net.named_parameters(remove_duplicate=False) if not prm.requires_grad}
    assert mine == frozen
