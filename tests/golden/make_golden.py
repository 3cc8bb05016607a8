"""Generate tests/golden/*.npz by running the UNMODIFIED reference on CPU.  Build container only.

    python tests/golden/make_golden.py [case ...]

For every case of cases.py: build the reference model (`/root/reference`, imported through
oracle/ref_import.py with third-party stubs), load the synthetic state_dict with strict=True (which
also proves that our parameter inventory matches the reference's names and shapes), run
`model(...)` in eval mode under no_grad, and store the reference's 14-tuple plus a few
intermediate activations captured with forward hooks.
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "comprehensive-transformer-tts_b200"))
sys.path.insert(0, HERE)

import cases  # noqa: E402
from oracle.ref_import import import_reference, reference_configs  # noqa: E402


def run_reference(name):
    c = cases.CASES[name]
    (p, m, t), sd, batch = cases.build_case(name)
    ref_model, _ = import_reference()
    rp, rm, rt = reference_configs(c["dataset"])
    rm["block_type"] = c["block_type"]
    rm["duration_modeling"]["learn_alignment"] = c["learn_alignment"]
    if c.get("prosody"):
        rm["prosody_modeling"]["model_type"] = c["prosody"]
    if c.get("pitch_type"):
        rp["preprocessing"]["pitch"]["pitch_type"] = c["pitch_type"]
    net = ref_model.CompTransTTS(rp, rm, rt).eval()
    net.load_state_dict(sd, strict=True)
    taps = {}
    net.encoder.register_forward_hook(lambda mod, i, o: taps.__setitem__("encoder_out", o[0].detach().clone()))
    net.decoder.register_forward_hook(lambda mod, i, o: taps.__setitem__("decoder_out", o[0].detach().clone()))
    net.decoder.register_forward_pre_hook(lambda mod, i: taps.__setitem__("decoder_in", i[0].detach().clone()))
    args, kw = cases.call_kwargs(batch)
    with torch.no_grad():
        out = net(*args, **kw)
    flat = {"ref." + k: v for k, v in cases.flatten_outputs(out).items()}
    for k, v in taps.items():
        flat["tap." + k] = v.numpy()[:, ::cases.TAP_STRIDE]  # every 4th frame keeps the fixtures small
    flat["in.texts"] = batch["texts"].numpy()
    flat["in.src_lens"] = batch["src_lens"].numpy()
    return flat


def main():
    names = sys.argv[1:] or list(cases.CASES)
    cwd = os.getcwd()
    for name in names:
        flat = run_reference(name)
        os.chdir(cwd)
        path = os.path.join(HERE, name + ".npz")
        np.savez_compressed(path, **flat)
        print("%s: %d arrays, %.1f KiB" % (name, len(flat), os.path.getsize(path) / 1024))


if __name__ == "__main__":
    main()
