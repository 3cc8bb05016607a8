"""Generate the training-step fixtures tests/golden/*_train.npz by running the reference on CPU.  Build container only.

    python tests/golden/make_golden_train.py [case ...]

The reference model is used as it is, in model.train() mode, with two process-level patches that its code needs to be
deterministic / runnable here (neither touches its arithmetic):
  * torch.nn.functional.dropout -> identity: the reference's dropout draws from torch's Philox stream, which no other
    implementation can reproduce, so training parity is DEFINED at dropout probability 0 (PostNet's 0.5 is hard-coded,
    modules.py:144-145, hence a patch rather than a config change);
  * torch.Tensor.cuda -> identity: coordconv.py:28,63 tests `torch.cuda.is_available` (the function object, always
    true) and moves tensors to the GPU unconditionally, which cannot work in a CPU-only container.
Stored per case: the 14-tuple of the forward, the value of cases.train_objective, GRAD_SAMPLES strided entries of its
gradient w.r.t. every parameter (plus each gradient's L2 norm and which parameters are frozen), and the BatchNorm
buffers after the step.
"""
import os
import sys

import numpy as np
import torch
import torch.nn.functional as F

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "comprehensive-transformer-tts_b200"))
sys.path.insert(0, HERE)

import cases  # noqa: E402
from oracle.ref_import import import_reference, reference_configs  # noqa: E402


def run_reference(name):
    c = cases.TRAIN_CASES[name]
    (p, m, t), sd, batch = cases.build_case(name)
    ref_model, _ = import_reference()
    rp, rm, rt = reference_configs(c["dataset"])
    rm["block_type"] = c["block_type"]
    rm["duration_modeling"]["learn_alignment"] = c["learn_alignment"]
    if c.get("prosody"):
        rm["prosody_modeling"]["model_type"] = c["prosody"]
    if c.get("pitch_type"):
        rp["preprocessing"]["pitch"]["pitch_type"] = c["pitch_type"]
    net = ref_model.CompTransTTS(rp, rm, rt)
    net.load_state_dict(sd, strict=True)
    net.train()
    args, kw = cases.call_kwargs(batch, c.get("step"))
    orig_dropout, orig_cuda = F.dropout, torch.Tensor.cuda
    F.dropout = lambda input, p=0.5, training=True, inplace=False: input
    torch.Tensor.cuda = lambda self, *a, **k: self
    try:
        out = net(*args, **kw)
        loss = cases.train_objective(out)
        loss.backward()
    finally:
        F.dropout, torch.Tensor.cuda = orig_dropout, orig_cuda
    flat = {"ref." + k: v for k, v in cases.flatten_outputs(out).items()}
    flat["loss"] = np.float64(loss.item())
    n_zero = 0
    for k, prm in net.named_parameters(remove_duplicate=False):
        g = prm.grad if prm.grad is not None else torch.zeros_like(prm)
        n_zero += int(prm.grad is None)
        gf = g.detach().reshape(-1)
        flat["grad." + k] = gf[torch.from_numpy(cases.grad_sample_index(gf.numel()))].numpy()
        flat["gnorm." + k] = np.float64(gf.double().norm().item())
        if not prm.requires_grad:
            flat["frozen." + k] = np.int8(1)      # nn.Parameter(requires_grad=False): sinusoid tables, bucket edges
    for k, buf in net.named_buffers():
        if k.endswith(("running_mean", "running_var", "num_batches_tracked")):
            flat["buf." + k] = buf.detach().numpy().copy()
    flat["in.texts"] = batch["texts"].numpy()
    return flat, n_zero


def main():
    names = sys.argv[1:] or list(cases.TRAIN_CASES)
    cwd = os.getcwd()
    for name in names:
        flat, n_zero = run_reference(name)
        os.chdir(cwd)
        path = os.path.join(HERE, name + ".npz")
        np.savez_compressed(path, **flat)
        print("%s: %d arrays (%d parameters without gradient), loss %.6f, %.1f KiB"
              % (name, len(flat), n_zero, flat["loss"], os.path.getsize(path) / 1024))


if __name__ == "__main__":
    main()
