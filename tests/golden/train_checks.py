"""Shared by the CPU (emulated kernels) and GPU training-step tests: run a TRAIN_CASES entry through the product module in
model.train() mode with dropout disabled, and compare with the reference's fixture."""
import numpy as np
import torch

import cases


def make_net(name, device):
    import ctts_b200
    (p, m, t), sd, batch = cases.build_case(name)
    net = ctts_b200.CompTransTTS(p, m, t)
    net.load_state_dict(sd, strict=True)
    net.to(device)
    net.train()
    return net, batch


def run_training_case(name, device, net=None, batch=None):
    if net is None:
        net, batch = make_net(name, device)
    args, kw = cases.call_kwargs(batch, cases.TRAIN_CASES[name].get("step"))

    def mv(v):
        if torch.is_tensor(v):
            return v.to(device)
        if isinstance(v, dict):
            return {k: mv(x) for k, x in v.items()}
        return v

    out = net(*[mv(a) for a in args], **{k: mv(v) for k, v in kw.items()})
    out_cpu = _to_cpu(out)
    loss = cases.train_objective(out_cpu)
    loss.backward()
    return net, out_cpu, loss


def _to_cpu(v):
    """Move the 14-tuple to the CPU keeping the autograd connection (the objective's cos weights live on the CPU)."""
    if torch.is_tensor(v):
        return v.cpu() if v.is_cuda else v
    if isinstance(v, dict):
        return {k: _to_cpu(x) for k, x in v.items()}
    if isinstance(v, (tuple, list)):
        return type(v)(_to_cpu(x) for x in v)
    return v


def check_against_fixture(net, out, loss, gold, out_atol=1e-4, out_rtol=2e-4, grad_scale=1.0, min_grads=150,
                          loss_rtol=1e-4, outlier_frac=0.005, outlier_bound=0.05, norm_rtol=1e-3, buf_atol=1e-5,
                          scalar_rtol=0.01, stats=None):
    flat = cases.flatten_outputs(out)
    n_out = 0
    report = []
    for key in gold.files:
        if key.startswith("ref."):
            k = key[4:]
            assert k in flat, "output lacks %s" % k
            a, b = gold[key], flat[k]
            assert a.shape == b.shape, (k, a.shape, b.shape)
            if a.dtype.kind in "biu":
                assert np.array_equal(a, b), k
            else:
                err = np.abs(b - a)
                tol = out_atol + out_rtol * np.abs(a)
                if stats is not None and a.size:
                    stats.setdefault("output_max_err", {})[k] = float(err.max())
                if not (err <= tol).all():
                    worst = (err - tol).argmax()
                    report.append("output %s: err %.3g (tol %.3g)" % (k, err.flat[worst], tol.flat[worst]))
            n_out += 1
    assert n_out >= 10
    if abs(loss.item() - float(gold["loss"])) > loss_rtol * max(1.0, abs(float(gold["loss"]))):
        report.append("loss %.6f vs %.6f" % (loss.item(), float(gold["loss"])))
    params = dict(net.named_parameters(remove_duplicate=False))
    n_grad = 0
    for key in gold.files:
        if not key.startswith("grad."):
            continue
        k = key[5:]
        prm = params[k]
        g = prm.grad
        gf = (g if g is not None else torch.zeros_like(prm)).detach().cpu().reshape(-1)
        mine = gf[torch.from_numpy(cases.grad_sample_index(gf.numel()))].numpy()
        ref = gold[key]
        norm = float(gold["gnorm." + k])
        tol = 2e-4 * max(norm / max(gf.numel(), 1) ** 0.5, 1e-6) + 1e-6
        err = np.abs(mine - ref)
        lim = grad_scale * (20 * tol + 2e-3 * np.abs(ref))
        # A ReLU (or bucket) whose pre-activation sits within float noise of zero flips between two float32
        # implementations with different summation order: that moves a handful of isolated gradient entries by a finite
        # amount while the norm stays put.  Allow <= 0.5 % of the sampled entries (at least 2) to be such outliers, each
        # bounded by 5 % of the tensor's rms; everything else must be inside the limit.
        n_bad = int((err > lim).sum())
        rms = norm / max(gf.numel(), 1) ** 0.5
        if stats is not None:
            stats["grad_worst_rel_rms"] = max(stats.get("grad_worst_rel_rms", 0.0), float(err.max() / max(rms, 1e-12))
                                              if norm > 1e-4 else 0.0)
            stats["grad_outlier_entries"] = stats.get("grad_outlier_entries", 0) + n_bad
            stats["grad_entries"] = stats.get("grad_entries", 0) + len(err)
        if gf.numel() == 1:
            # a scalar gradient (learnable positional scale) is one long sum with heavy cancellation: its error scales
            # with the sum of the magnitudes of its terms, not with its own value
            if abs(float(mine[0]) - float(ref[0])) > scalar_rtol * max(abs(float(ref[0])), 1e-3):
                report.append("scalar grad %s: %.6g vs %.6g" % (k, float(mine[0]), float(ref[0])))
            n_grad += 1
            continue
        if n_bad > max(2, int(len(err) * outlier_frac)) or (n_bad and err.max() > outlier_bound * rms + 1e-6):
            report.append("grad %s: %d entries over the limit, max err %.3g (limit %.3g, rms %.3g)" % (
                k, n_bad, err.max(), lim.flat[err.argmax()], rms))
        elif abs(float(gf.double().norm()) - norm) > grad_scale * (norm_rtol * norm + 1e-6):
            report.append("grad norm %s: %.6g vs %.6g" % (k, float(gf.double().norm()), norm))
        n_grad += 1
    assert n_grad >= min_grads
    bufs = dict(net.named_buffers())
    for key in gold.files:
        if key.startswith("buf.") and key[4:] in bufs:
            a, b = gold[key], bufs[key[4:]].detach().cpu().numpy()
            if not np.allclose(b, a, atol=buf_atol, rtol=1e-4):
                report.append("buffer %s: max err %.3g" % (key[4:], np.abs(b - a).max()))
    assert not report, "\n".join(report[:40]) + ("\n... %d more" % (len(report) - 40) if len(report) > 40 else "")
