"""Where does the gradient difference against the fp32 CPU oracle come from?

    python tools/grad_gate.py [--n 8] [--size 224] [--out profiles/r2_grad_gate.md]

One RSP pretext step (TripletNet + Classifier(768,6), CE, pretrain_BreastPathQ.py:53-60) on the
same seeded weights and inputs, run by several implementations; every one is compared with the
plain fp32 CPU oracle (oracle/ref_net.py, the reference's own arithmetic):

  cpu_fp64        the oracle in float64            -> how far fp32 itself is from the exact result
  cpu_tf32_bwd    fp32 forward, TF32 backward ops   -> operand-exact model of libb2n's arithmetic
  cpu_tf32_all    TF32 operands everywhere          -> torch's GPU default (cudnn.allow_tf32=True)
  gpu_cudnn_fp32  oracle modules on the GPU, allow_tf32=False     (GPU only)
  gpu_cudnn_tf32  oracle modules on the GPU, allow_tf32=True      (GPU only)
  b2n             this repo's CUDA path                            (GPU only)
  b2n_precise     this repo with error-compensated backward convs  (GPU only, if available)

Reported per implementation: logits max-rel, loss rel, and the per-tensor gradient relative L2
(worst, median, and the FLOP-heavy conv weights' worst).  Test infrastructure: uses oracle/.
"""
import argparse
import json
import os
import statistics
import sys

import torch
import torch.nn.functional as F

ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, ROOT)
from oracle import ref_net as O  # noqa: E402


def run_oracle(st, hs, inputs, target, dev="cpu", dtype=torch.float32, mode=None):
    O.TF32_OPERANDS, O.TF32_BACKWARD = mode == "all", mode == "bwd"
    try:
        m, c = O.TripletNet("resnet18"), O.Classifier(768, 6)
        m.load_state_dict(st); c.load_state_dict(hs)
        m, c = m.to(dev, dtype).train(), c.to(dev, dtype).train()
        out = c(m(*[i.to(dev, dtype) for i in inputs]))
        loss = F.cross_entropy(out, target.to(dev))
        loss.backward()
        grads = {n: p.grad.detach().double().cpu() for mod in (m, c) for n, p in mod.named_parameters()}
        return out.detach().double().cpu(), float(loss.detach()), grads
    finally:
        O.TF32_OPERANDS = O.TF32_BACKWARD = False


def run_b2n(st, hs, inputs, target, precise=False):
    import ssl_cr_histo_b200.net as net
    from ssl_cr_histo_b200 import trunk
    old = getattr(trunk, "PRECISE_BACKWARD", None)
    if precise:
        if old is None:
            return None
        trunk.PRECISE_BACKWARD = True
    try:
        m, c = net.TripletNet("resnet18"), net.Classifier(768, 6)
        m.load_state_dict(st); c.load_state_dict(hs)
        m, c = m.cuda().train(), c.cuda().train()
        out = c(m(*[i.cuda() for i in inputs]))
        loss = F.cross_entropy(out, target.cuda())
        loss.backward()
        grads = {n: p.grad.detach().double().cpu() for mod in (m, c) for n, p in mod.named_parameters()}
        return out.detach().double().cpu(), float(loss.detach()), grads
    finally:
        if old is not None:
            trunk.PRECISE_BACKWARD = old


def compare(name, res, base):
    out, loss, grads = res
    bout, bloss, bgrads = base
    rels = {}
    for k, g in grads.items():
        rels[k] = float((g - bgrads[k]).norm() / bgrads[k].norm().clamp_min(1e-300))
    convs = [v for k, v in rels.items() if k.endswith("conv1.weight") or k.endswith("conv2.weight")
             or "downsample.0" in k]
    worst = max(rels, key=rels.get)
    return {"impl": name, "logits_max_rel": float((out - bout).abs().max() / bout.abs().max()),
            "loss_rel": abs(loss - bloss) / abs(bloss), "grad_rel_l2_worst": rels[worst],
            "worst_tensor": worst, "grad_rel_l2_median": statistics.median(rels.values()),
            "conv_weight_grad_rel_l2_worst": max(convs),
            "conv_weight_grad_rel_l2_median": statistics.median(convs)}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--n", type=int, default=8)
    ap.add_argument("--size", type=int, default=224)
    ap.add_argument("--out", default=None)
    ap.add_argument("--skip-fp64", action="store_true")
    a = ap.parse_args()
    torch.set_num_threads(os.cpu_count() or 1)
    st, hs = O.reference_state(42, ("classifier", 6))
    inputs = [O.synthetic_patches(a.n, a.size, seed=s) for s in (0, 1, 2)]
    target = torch.randint(0, 6, (a.n,), generator=torch.Generator().manual_seed(5))
    base = run_oracle(st, hs, inputs, target)
    rows = []
    if not a.skip_fp64:
        rows.append(compare("cpu_fp64", run_oracle(st, hs, inputs, target, dtype=torch.float64), base))
    rows.append(compare("cpu_tf32_bwd", run_oracle(st, hs, inputs, target, mode="bwd"), base))
    rows.append(compare("cpu_tf32_all", run_oracle(st, hs, inputs, target, mode="all"), base))
    if torch.cuda.is_available():
        for allow in (False, True):
            torch.backends.cudnn.allow_tf32 = allow
            torch.backends.cuda.matmul.allow_tf32 = False
            rows.append(compare("gpu_cudnn_%s" % ("tf32" if allow else "fp32"),
                                run_oracle(st, hs, inputs, target, dev="cuda"), base))
        rows.append(compare("b2n", run_b2n(st, hs, inputs, target), base))
        r = run_b2n(st, hs, inputs, target, precise=True)
        if r is not None:
            rows.append(compare("b2n_precise", r, base))
    hdr = ("| implementation | logits max-rel | loss rel | grad rel-L2 worst (tensor) | grad rel-L2 median | "
           "conv dW rel-L2 worst | conv dW rel-L2 median |\n|---|---:|---:|---:|---:|---:|---:|\n")
    body = "".join("| %s | %.2e | %.2e | %.2e (%s) | %.2e | %.2e | %.2e |\n" % (
        r["impl"], r["logits_max_rel"], r["loss_rel"], r["grad_rel_l2_worst"], r["worst_tensor"],
        r["grad_rel_l2_median"], r["conv_weight_grad_rel_l2_worst"], r["conv_weight_grad_rel_l2_median"])
        for r in rows)
    text = ("RSP pretext step, N=%d triples at %dx%d, seed-42 init; every row vs the fp32 CPU oracle\n\n"
            % (a.n, a.size, a.size)) + hdr + body
    print(text)
    print(json.dumps(rows))
    if a.out:
        with open(a.out, "w") as f:
            f.write(text)


if __name__ == "__main__":
    main()
