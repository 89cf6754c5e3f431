"""Layer-by-layer bring-up check of the CUDA trunk against the CPU oracle (run on a GPU box).

    python tools/gpu_check.py [N] [size]

Prints max-abs-diff / max-abs-ref for every intermediate tensor of a train-mode forward pass,
then for every parameter gradient and BN buffer after one backward pass.
"""
import os
import sys

import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))

import ssl_cr_histo_b200.net as net  # noqa: E402
from ssl_cr_histo_b200.trunk import _TrunkFn  # noqa: E402
from oracle import ref_net  # noqa: E402


def rel(a, b):
    a, b = a.double().cpu(), b.double().cpu()
    return float(((a - b).abs().max() / b.abs().max().clamp_min(1e-30)).detach())


def nhwc(t):
    return t.permute(0, 3, 1, 2)


def main():
    N = int(sys.argv[1]) if len(sys.argv) > 1 else 4
    size = int(sys.argv[2]) if len(sys.argv) > 2 else 64
    torch.manual_seed(42)
    ref = ref_net.TripletNet_Finetune("resnet18")
    mine = net.TripletNet_Finetune("resnet18")
    mine.load_state_dict(ref.state_dict())
    mine = mine.cuda()
    x = ref_net.synthetic_patches(N, size, seed=0)

    # ---- oracle intermediates
    inter = {}

    def hook(name):
        def f(_m, _i, o):
            inter[name] = o.detach()
        return f

    tr = ref.model
    tr.conv1.register_forward_hook(hook("y0"))
    for li in (1, 2, 3, 4):
        for bi, blk in enumerate(getattr(tr, "layer%d" % li)):
            p = "l%d.%d." % (li, bi)
            blk.conv1.register_forward_hook(hook(p + "y1"))
            blk.conv2.register_forward_hook(hook(p + "y2"))
            blk.register_forward_hook(hook(p + "a_out"))
            if blk.downsample is not None:
                blk.downsample[0].register_forward_hook(hook(p + "yd"))
    ref.train()
    e_ref = tr(x)

    class Ctx:
        needs_input_grad = (False,) * 4 + (True,) * 60

    ctx = Ctx()
    trunk = mine.model.train()
    params = list(trunk.parameters())
    with torch.enable_grad():
        e = _TrunkFn.forward(ctx, x.cuda(), trunk, 1, True, *params)
    torch.cuda.synchronize()
    sv = ctx.saved
    print("%-14s %s" % ("tensor", "max|d|/max|ref|"))
    print("%-14s %.3e" % ("y0", rel(nhwc(sv["y0"]), inter["y0"])))
    names = ["l%d.%d." % (li, bi) for li in (1, 2, 3, 4) for bi in (0, 1)]
    for p, rec in zip(names, sv["blocks"]):
        for k in ("y1", "y2", "yd", "a_out"):
            if k in rec:
                print("%-14s %.3e" % (p + k, rel(nhwc(rec[k]), inter[p + k])))
    print("%-14s %.3e" % ("features", rel(e, e_ref)))
    for name in ("bn1.running_mean", "bn1.running_var", "layer4.1.bn2.running_mean",
                 "layer4.1.bn2.running_var", "layer2.0.downsample.1.running_var"):
        print("%-34s %.3e" % (name, rel(trunk.state_dict()[name], tr.state_dict()[name])))

    # ---- full step through the public modules: grads + buffers
    torch.manual_seed(42)
    ref = ref_net.TripletNet_Finetune("resnet18")
    cls_ref = ref_net.FinetuneResNet(9)
    mine = net.TripletNet_Finetune("resnet18")
    cls = net.FinetuneResNet(9)
    mine.load_state_dict(ref.state_dict())
    cls.load_state_dict(cls_ref.state_dict())
    mine, cls = mine.cuda(), cls.cuda()
    target = torch.randint(0, 9, (N,), generator=torch.Generator().manual_seed(1))
    ref.train(); mine.train()
    out_ref = cls_ref(ref(x))
    loss_ref = torch.nn.functional.cross_entropy(out_ref, target)
    loss_ref.backward()
    out = cls(mine(x.cuda()))
    loss = torch.nn.functional.cross_entropy(out, target.cuda())
    loss.backward()
    torch.cuda.synchronize()
    print("logits %.3e  loss %.6f vs %.6f" % (rel(out, out_ref), loss.item(), loss_ref.item()))
    worst = 0.0
    for (n1, p1), (n2, p2) in zip(mine.named_parameters(), ref.named_parameters()):
        g1, g2 = p1.grad.double().cpu(), p2.grad.double()
        r = float((g1 - g2).norm() / g2.norm().clamp_min(1e-30))
        worst = max(worst, r)
        print("grad %-36s relL2 %.3e" % (n1, r))
    print("worst grad relL2 %.3e" % worst)
    bw = 0.0
    for (n1, b1), (n2, b2) in zip(mine.named_buffers(), ref.named_buffers()):
        r = rel(b1.float(), b2.float())
        bw = max(bw, r)
    print("worst buffer rel %.3e" % bw)


if __name__ == "__main__":
    main()
