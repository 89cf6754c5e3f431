"""Debug: fused vs separate 1x1 shortcut gradient in the merged stride-2 data gradient."""
import copy, os, sys
import torch
import torch.nn.functional as F
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests"))
from ssl_cr_histo_b200 import trunk
from ssl_cr_histo_b200._lib import call
from oracle import ref_net as O
from test_gpu_parity import pair

DEV = "cuda"
_, _, gm, gh = pair("finetune", ("finetune", 9))
x = O.synthetic_patches(5, 96, seed=120).to(DEV)
target = torch.tensor([0, 3, 8, 1, 5], device=DEV)

def grads():
    m, c = copy.deepcopy(gm).train(), copy.deepcopy(gh)
    F.cross_entropy(c(m(x)), target).backward()
    return [p.grad.clone() for p in m.parameters()]

names = [n for n, _ in gm.named_parameters()]
f1, f2 = grads(), grads()
print("fused repeatable:", all(torch.equal(a, b) for a, b in zip(f1, f2)))
trunk.FUSED_S2_SHORTCUT = False
u1, u2 = grads(), grads()
print("unfused repeatable:", all(torch.equal(a, b) for a, b in zip(u1, u2)))
for n, a, b in zip(names, u1, f1):
    r = float((a - b).norm() / a.norm().clamp_min(1e-30))
    if "conv" in n or "downsample.0" in n:
        print("%-40s %.3e" % (n, r))

# direct kernel comparison on real-valued TF32-rounded data
def tf32(t):
    return O.tf32_round(t)
g = torch.Generator().manual_seed(1)
for (N, H, W, Cin, Cout) in ((5, 24, 24, 64, 128), (5, 12, 12, 128, 256), (5, 6, 6, 256, 512)):
    P, Q = (H + 1) // 2, (W + 1) // 2
    w3 = (torch.randn(Cout, Cin, 3, 3, generator=g) * 0.05).to(DEV)
    w1 = (torch.randn(Cout, Cin, 1, 1, generator=g) * 0.05).to(DEV)
    dy3 = tf32(torch.randn(N, P, Q, Cout, generator=g)).to(DEV)
    dy1 = tf32(torch.randn(N, P, Q, Cout, generator=g)).to(DEV)
    wm = torch.empty(Cin, 9 * Cout, device=DEV); call("b2n_pack_weight_dgrad_s2m", w3, wm, Cout, Cin)
    wd = torch.empty(Cin, Cout, device=DEV); call("b2n_pack_weight_dgrad", w1, wd, Cout, Cin, 1, 1)
    a = torch.full((N, H, W, Cin), float("nan"), device=DEV)
    b = torch.full((N, H, W, Cin), float("nan"), device=DEV)
    call("b2n_conv_fwd", dy1, None, None, wd, None, None, a, None, None, N, P, Q, Cout, Cin, 1, 1,
         1, 0, 0, 0, 0, None, None, None, None, None, None, 0, 0, None, None, 2, 0, 0, H, W,
         None, None, None, None, None, None)
    call("b2n_conv_dgrad_s2", dy3, wm, a, N, P, Q, Cout, Cin, H, W, a, None)
    call("b2n_conv_dgrad_s2_sc", dy3, wm, dy1, wd, b, N, P, Q, Cout, Cin, H, W, None)
    print((N, H, W, Cin, Cout), "kernel rel diff %.3e  max abs %.3e" % (float((a - b).norm() / a.norm()), float((a - b).abs().max())))
