"""Multi-GPU check of the backward pass' stream handling (torchrun --nproc-per-node 2 tools/ddp_check.py):
gradients of one consistency-style step through ddp.GradAllReducer with overlapped buckets must be
bit-identical whether the weight gradients run on the main stream or, ordered, on the side stream
(trunk.OVERLAP_WGRAD = 0 / 2), eagerly and replayed from a CUDA graph."""
import copy
import os
import sys

import torch
import torch.distributed as dist
import torch.nn.functional as F

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import ssl_cr_histo_b200.net as net  # noqa: E402
from ssl_cr_histo_b200 import ddp, trunk  # noqa: E402

rank, local = int(os.environ.get("RANK", "0")), int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
torch.manual_seed(42)
base = net.TripletNet_Finetune("resnet18").to(dev).train()
head = net.FinetuneResNet(9).to(dev).train()
x = torch.randint(0, 256, (24, 3, 96, 96), generator=torch.Generator().manual_seed(100 + rank)).float().to(dev)
t = torch.randint(0, 9, (24,), generator=torch.Generator().manual_seed(7 + rank)).to(dev)


def run(mode, steps=3):
    trunk.OVERLAP_WGRAD = mode
    m, h = copy.deepcopy(base), copy.deepcopy(head)
    params = list(m.parameters()) + list(h.parameters())
    red = ddp.GradAllReducer(params, overlap=True)
    out = []
    for _ in range(steps):
        red.zero_grad()
        F.cross_entropy(h(m(x)), t).backward()
        red.all_reduce(average=True)
        torch.cuda.synchronize()
        out.append(torch.cat([p.grad.flatten() for p in params]).clone())
    return out


ref = run(0)
for mode in (2, 1):
    got = run(mode)
    same = all(torch.equal(a, b) for a, b in zip(ref, got))
    worst = max(float((a - b).abs().max()) for a, b in zip(ref, got))
    flag = torch.tensor([1 if same else 0], device=dev)
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    if rank == 0:
        print("OVERLAP_WGRAD=%d: gradients %s (max |diff| %.3e)" % (mode, "bit-identical" if int(flag) else "DIFFER", worst))
dist.destroy_process_group()
