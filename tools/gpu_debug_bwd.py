"""Targeted backward-pass debugging (GPU box): heads alone, then trunk alone with per-unit
input-gradient comparison against the CPU oracle."""
import os
import sys

import torch
import torch.nn.functional as F

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))

import ssl_cr_histo_b200.net as net  # noqa: E402
from oracle import ref_net  # noqa: E402


def rl2(a, b):
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    return float((a - b).norm() / b.norm().clamp_min(1e-30))


def heads_only(N=8):
    torch.manual_seed(1)
    ref = ref_net.TripletNet_Finetune("resnet18")
    mine = net.TripletNet_Finetune("resnet18")
    mine.load_state_dict(ref.state_dict())
    mine = mine.cuda()
    x = torch.randn(N, 1024)
    dy = torch.randn(N, 256)
    xr = x.clone().requires_grad_(True)
    xm = x.clone().cuda().requires_grad_(True)
    yr = ref.fc(xr)
    yr.backward(dy)
    from ssl_cr_histo_b200 import heads
    ym = heads.mlp2(xm, mine.fc[0], mine.fc[2])
    ym.backward(dy.cuda())
    print("mlp2 fwd %.2e dx %.2e" % (rl2(ym, yr), rl2(xm.grad, xr.grad)))
    for (n1, p1), (n2, p2) in zip(mine.fc.named_parameters(), ref.fc.named_parameters()):
        print("  mlp2 grad %-10s %.2e" % (n1, rl2(p1.grad, p2.grad)))


def trunk_only(N=4, size=64):
    torch.manual_seed(42)
    ref = ref_net.ResNet18Trunk().train()
    mine = net.TripletNet("resnet18").model
    mine.load_state_dict(ref.state_dict())
    mine = mine.cuda().train()
    x = ref_net.synthetic_patches(N, size, seed=0)
    R = torch.randn(N, 512, generator=torch.Generator().manual_seed(3))

    # capture oracle gradients w.r.t. every block input / output
    grads = {}

    def keep(name):
        def f(g):
            grads[name] = g.detach()
        return f

    def fwd_hook(name):
        def f(_m, inp, out):
            if out.requires_grad:
                out.register_hook(keep(name))
        return f

    ref.bn1.register_forward_hook(fwd_hook("bn0_out"))
    for li in (1, 2, 3, 4):
        for bi, blk in enumerate(getattr(ref, "layer%d" % li)):
            blk.register_forward_hook(fwd_hook("l%d.%d.out" % (li, bi)))
            blk.conv1.register_forward_hook(fwd_hook("l%d.%d.y1" % (li, bi)))
            blk.conv2.register_forward_hook(fwd_hook("l%d.%d.y2" % (li, bi)))
    e_ref = ref(x)
    (e_ref * R).sum().backward()
    e = mine(x.cuda())
    # monkey-patch: record the tensors flowing through bn_backward / _conv in my backward
    import ssl_cr_histo_b200.trunk as T
    trace = []
    orig_call = T.call

    def spy(name, *args, **kw):
        orig_call(name, *args, **kw)
        if name == "b2n_bn_bwd_apply":
            trace.append(("dy", args[9].clone(), args[0].clone()))
        if name == "b2n_pool_bn_bwd_apply":
            trace.append(("dy", args[9].clone(), args[0].clone()))
    T.call = spy
    (e * R.cuda()).sum().backward()
    T.call = orig_call
    torch.cuda.synchronize()
    print("features %.2e" % rl2(e, e_ref))
    # trace order: l4.1 bn2, l4.1 bn1, l4.0 bn2, l4.0 bn1, l4.0 bnd, ...
    order = []
    for li in (4, 3, 2, 1):
        for bi in (1, 0):
            order += ["l%d.%d.y2" % (li, bi), "l%d.%d.y1" % (li, bi)]
            if bi == 0 and li > 1:
                order.append("l%d.%d.yd" % (li, bi))
    order.append("y0")
    for name, (_, dy, g) in zip(order, trace):
        if name in grads:
            print("d/d %-10s relL2 %.2e   (|g_in| %.3e)" % (name, rl2(dy.permute(0, 3, 1, 2), grads[name]),
                                                        float(g.norm())))
    worst = 0
    for (n1, p1), (n2, p2) in zip(mine.named_parameters(), ref.named_parameters()):
        r = rl2(p1.grad, p2.grad)
        worst = max(worst, r)
        print("grad %-34s %.2e" % (n1, r))
    print("worst %.2e" % worst)


if __name__ == "__main__":
    heads_only()
    trunk_only()
