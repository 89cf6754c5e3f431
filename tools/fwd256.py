"""The north_star's target shape: fused conv3x3 + BN + ReLU forward at batch 256 x 224 x 224 --
one eval-mode (BatchNorm folded into the conv epilogue) trunk pass of the drop-in model, for ncu.

    ncu --metrics ... python tools/fwd256.py [batch] [train]

With `train` the pass runs in training mode (batch statistics in the conv epilogues, separate
BatchNorm-apply kernels) -- the student's forward of the bench step.
"""
import os
import sys

import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import ssl_cr_histo_b200.net as net  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 256
torch.manual_seed(42)
m = net.TripletNet_Finetune("resnet18").cuda()
m = m.train() if len(sys.argv) > 2 and sys.argv[2] == "train" else m.eval()
x = torch.randint(0, 256, (n, 3, 224, 224), dtype=torch.uint8, generator=torch.Generator().manual_seed(0)).float().cuda()
with torch.no_grad():
    for _ in range(3):          # two warm-up passes, the third is the one to read
        f = m(x)
torch.cuda.synchronize()
print("features", tuple(f.shape), float(f.abs().mean()))
