"""CPU tests: the oracle restatement (oracle/ref_net.py) against the golden vectors produced by
the REAL reference modules (oracle/make_golden.py), against the installed torchvision trunk,
and -- when /root/reference is mounted -- against the reference classes directly."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.nn.functional as F

from oracle import ref_net as O
from util import golden, max_rel

torch.set_num_threads(min(8, os.cpu_count() or 1))
TOL = 2e-5  # same fp32 math, only summation order inside the CPU conv kernels may differ


def _load(model, head, head_spec):
    st, hs = O.reference_state(42, head_spec)
    model.load_state_dict(st)
    head.load_state_dict(hs)


def _buffers_fp(m):
    return np.array([float(b.double().sum()) for b in m.buffers()])


def _grad_norms(mods):
    return np.array([float(p.grad.double().norm()) for m in mods for p in m.parameters()])


def test_init_matches_reference_fingerprint():
    g = golden("init_seed42.npz")
    st, hs = O.reference_state(42, ("classifier", 6))
    assert list(g["keys"]) == list(st.keys())
    for row, (k, v) in zip(g["fp"], st.items()):
        f = v.double().flatten()
        assert abs(float(f.sum()) - row[0]) <= 1e-9 * max(1.0, abs(row[0])), k
        assert abs(float(f.abs().sum()) - row[1]) <= 1e-9 * max(1.0, abs(row[1])), k
        assert np.array_equal(f[:4].numpy(), row[2:2 + min(4, f.numel())]), k
    m = O.TripletNet("resnet18")
    assert [n for n, _ in m.named_parameters()] == list(g["param_names"])
    assert len(list(m.named_parameters())) == 64 and len(list(m.buffers())) == 60


def test_trunk_equals_torchvision():
    import torchvision

    torch.manual_seed(3)
    tv = torchvision.models.resnet18(weights=None)
    tv.fc = torch.nn.Sequential()
    mine = O.ResNet18Trunk()
    mine.load_state_dict(tv.state_dict())
    x = O.synthetic_patches(2, 96, seed=4)
    tv.train(); mine.train()
    assert max_rel(mine(x), tv(x)) < TOL
    for (k1, b1), (k2, b2) in zip(mine.named_buffers(), tv.named_buffers()):
        assert k1 == k2 and torch.allclose(b1.float(), b2.float(), rtol=1e-5, atol=1e-6), k1
    tv.eval(); mine.eval()
    assert max_rel(mine(x), tv(x)) < TOL


def test_cfg1_rsp_forward_golden():
    g = golden("cfg1_rsp_forward.npz")
    model, cls = O.TripletNet("resnet18"), O.Classifier(768, 6)
    _load(model, cls, ("classifier", 6))
    model.train(); cls.train()
    i1, i2, i3 = (O.synthetic_patches(8, 224, seed=s) for s in (0, 1, 2))
    feats, logits = [], []
    with torch.no_grad():
        for b in range(4):
            sl = slice(2 * b, 2 * b + 2)
            f = model(i1[sl], i2[sl], i3[sl])
            feats.append(f); logits.append(cls(f))
    feats, logits = torch.cat(feats), torch.cat(logits)
    assert max_rel(feats, g["feats"]) < TOL
    assert max_rel(logits, g["logits"]) < TOL
    assert np.array_equal(torch.argmax(logits, 1).numpy(), g["pred"])
    assert np.allclose(_buffers_fp(model), g["buffers"], rtol=1e-5)


def test_rsp_step_golden():
    g = golden("rsp_step.npz")
    model, cls = O.TripletNet("resnet18"), O.Classifier(768, 6)
    _load(model, cls, ("classifier", 6))
    model.train(); cls.train()
    opt = O.make_rsp_optimizer(list(model.parameters()) + list(cls.parameters()))
    i1, i2, i3 = (O.synthetic_patches(2, 64, seed=s) for s in (0, 1, 2))
    out = O.rsp_pretrain_step(model, cls, opt, i1, i2, i3, torch.tensor([3, 5]))
    assert abs(float(out["loss"]) - float(g["loss"])) < 1e-5
    assert max_rel(out["output"], g["output"]) < TOL
    assert np.array_equal(out["pred"].numpy(), g["pred"])
    assert np.allclose(_grad_norms([model, cls]), g["grad_norms"], rtol=2e-3, atol=1e-7)
    assert np.allclose(_buffers_fp(model), g["buffers"], rtol=1e-5)
    pa = np.array([float(p.detach().double().sum()) for m in (model, cls) for p in m.parameters()])
    assert np.allclose(pa, g["params_after"], rtol=1e-4, atol=1e-4)


@pytest.mark.parametrize("kind", ["mse", "ce"])
def test_cr_step_golden(kind):
    g = golden("cr_step_%s.npz" % kind)
    C = 1 if kind == "mse" else 9
    student, cls_s = O.TripletNet_Finetune("resnet18"), O.FinetuneResNet(C)
    _load(student, cls_s, ("finetune", C))
    teacher, cls_t = O.teacher_handoff(student), O.teacher_handoff(cls_s)
    O.freeze_by_index(teacher, 64)
    for p in cls_t.parameters():
        p.requires_grad = False
    teacher.eval(); cls_t.eval(); student.train(); cls_s.train()
    opt = O.make_cr_optimizer(list(student.parameters()) + list(cls_s.parameters()))
    tx = torch.tensor([0.25] * 3) if kind == "mse" else torch.tensor([4, 4, 4])
    out = O.consistency_step(teacher, student, cls_t, cls_s, opt, O.synthetic_patches(3, 64, seed=10),
                             tx, O.synthetic_patches(2, 64, seed=11),
                             O.synthetic_patches(2, 64, seed=12), 1.0, kind)
    for k in ("sup", "cons"):
        assert abs(float(out[k]) - float(g[k])) < 1e-5 * max(1.0, abs(float(g[k]))), k
    assert abs(float(out["loss"]) - float(g["final"])) < 1e-5 * max(1.0, abs(float(g["final"])))
    for k in ("logits_x", "logits_u_s", "logits_u_w"):
        assert max_rel(out[k], g[k]) < 5e-5, k
    # the reference's three trunk calls bump every BN counter by 3 (models/net.py:88-90)
    assert int(student.model.bn1.num_batches_tracked) == 3
    assert np.allclose(_buffers_fp(student), g["buffers"], rtol=1e-5)
    assert np.allclose(_grad_norms([student, cls_s]), g["grad_norms"], rtol=5e-3, atol=1e-7)


def test_finetune_step_golden():
    g = golden("finetune_step.npz")
    model, cls = O.TripletNet_Finetune("resnet18"), O.FinetuneResNet(9)
    _load(model, cls, ("finetune", 9))
    model.train(); cls.train()
    opt = torch.optim.Adam(list(model.parameters()) + list(cls.parameters()), lr=1e-5,
                           betas=(0.9, 0.999), weight_decay=1e-4)
    out = O.finetune_step(model, cls, opt, O.synthetic_patches(4, 64, seed=20),
                          torch.tensor([0, 8, 3, 3]))
    assert abs(float(out["loss"]) - float(g["loss"])) < 1e-5
    assert max_rel(out["output"], g["output"]) < TOL
    assert np.array_equal(out["pred"].numpy(), g["pred"])
    assert int(model.model.bn1.num_batches_tracked) == int(g["nbt"][0]) == 3
    assert np.allclose(_grad_norms([model, cls]), g["grad_norms"], rtol=5e-3, atol=1e-7)


def test_lookahead_golden():
    g = golden("lookahead.npz")
    shapes = [(7,), (3, 5), (2, 2, 2)]
    flat, grads = torch.tensor(g["init"]), torch.tensor(g["grads"])
    params, gs, off = [], [], 0
    for s in shapes:
        n = int(np.prod(s))
        params.append(flat[off:off + n].clone().view(s)); gs.append(grads[off:off + n].view(s)); off += n
    cached = [p.clone() for p in params]
    for step in range(7):
        for p, gr in zip(params, gs):
            p.add_(gr, alpha=-0.1)            # inner SGD(lr=0.1) step (lookahead.py:87)
        if (step + 1) % 5 == 0:
            O.lookahead_pull([torch.nn.Parameter(p) for p in params], cached, 0.5)
        now = torch.cat([p.flatten() for p in params])
        assert torch.allclose(now, torch.tensor(g["trace"][step]), rtol=1e-6, atol=1e-7), step


def test_tf32_round_is_rna():
    x = torch.tensor([1.0, 1.0 + 2 ** -11, 1.0 + 2 ** -11 + 2 ** -20, -3.1415926, 255.0, 0.0])
    r = O.tf32_round(x)
    assert r[0] == 1.0 and r[1] == 1.0 + 2 ** -10 and r[2] == 1.0 + 2 ** -10
    assert r[4] == 255.0 and r[5] == 0.0
    assert abs(float(r[3]) + 3.1415926) <= 3.1415926 * 2 ** -11


@pytest.mark.skipif(not os.path.isdir("/root/reference/models"), reason="reference tree absent")
def test_against_live_reference_modules():
    import contextlib
    import io

    sys.dont_write_bytecode = True
    sys.path.insert(0, "/root/reference")
    try:
        with contextlib.redirect_stdout(io.StringIO()):
            import models.net as refnet
            torch.manual_seed(7)
            ref = refnet.TripletNet_Finetune("resnet18")
            ref_cls = refnet.FinetuneResNet(2)
    finally:
        sys.path.remove("/root/reference")
    mine, mine_cls = O.TripletNet_Finetune("resnet18"), O.FinetuneResNet(2)
    mine.load_state_dict(ref.state_dict()); mine_cls.load_state_dict(ref_cls.state_dict())
    x = O.synthetic_patches(3, 64, seed=9)
    ref.train(); mine.train()
    a, b = ref_cls(ref(x)), mine_cls(mine(x))
    assert max_rel(b, a) < TOL
    F.cross_entropy(a, torch.tensor([0, 1, 1])).backward()
    F.cross_entropy(b, torch.tensor([0, 1, 1])).backward()
    for (n1, p1), (n2, p2) in zip(ref.named_parameters(), mine.named_parameters()):
        assert n1 == n2
        assert torch.allclose(p1.grad, p2.grad, rtol=2e-3, atol=1e-6), n1
