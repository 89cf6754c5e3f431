"""Generates tests/golden/*.npz by running the REAL reference modules (imported unmodified from
/root/reference: models/net.py, models/optimiser/RAdam/lookahead.py) on seeded synthetic inputs.

Run in the build container only (the reference tree does not exist on the GPU box):

    PYTHONDONTWRITEBYTECODE=1 python oracle/make_golden.py

The fixtures pin the oracle restatement (oracle/ref_net.py): tests/test_oracle.py recomputes
every quantity with the restated classes / step bodies and compares.  Loop bodies are restated
from the driver scripts (which cannot be imported: they need openslide/albumentations/h5py ...,
and eval_Kather_SSL.py:243 is a SyntaxError) -- see oracle/ref_net.py for the line references.
"""
import contextlib
import io
import os
import sys

import numpy as np
import torch
import torch.nn.functional as F

REF = "/root/reference"
HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.path.join(HERE, "..", "tests", "golden")
sys.dont_write_bytecode = True
sys.path.insert(0, os.path.join(HERE, ".."))
sys.path.insert(0, REF)

with contextlib.redirect_stdout(io.StringIO()):  # the constructors print the whole model
    import models.net as refnet  # noqa: E402
from models.optimiser.RAdam.lookahead import Lookahead  # noqa: E402
from oracle import ref_net as O  # noqa: E402  (only for the shared input generators)


def quiet(fn, *a):
    with contextlib.redirect_stdout(io.StringIO()):
        return fn(*a)


def np_(t):
    return t.detach().cpu().numpy()


def state_fingerprint(sd):
    """sum, abs-sum and first 4 values of every tensor -- enough to pin an init bit-for-bit."""
    rows = []
    for k, v in sd.items():
        f = v.double().flatten()
        head = torch.zeros(4, dtype=torch.float64)
        head[:min(4, f.numel())] = f[:4]
        rows.append(torch.cat([f.sum().view(1), f.abs().sum().view(1), head]))
    return torch.stack(rows).numpy(), np.array(list(sd.keys()))


def grad_norms(mods):
    return np.array([float(p.grad.double().norm()) if p.grad is not None else -1.0
                     for m in mods for p in m.parameters()])


def buffers_fp(mods):
    return np.array([float(b.double().sum()) for m in mods for b in m.buffers()])


def params_fp(mods):
    return np.array([float(p.detach().double().sum()) for m in mods for p in m.parameters()])


def golden_init():
    torch.manual_seed(42)
    model = quiet(refnet.TripletNet, "resnet18")
    cls = quiet(refnet.Classifier, 768, 6)
    fp, keys = state_fingerprint(model.state_dict())
    fpc, keysc = state_fingerprint(cls.state_dict())
    np.savez(os.path.join(OUT, "init_seed42.npz"), fp=fp, keys=keys, fp_cls=fpc, keys_cls=keysc,
             param_names=np.array([n for n, _ in model.named_parameters()]))


def golden_cfg1():
    """BASELINE.json configs[0]: RSP pretext forward, 8 synthetic 224x224 triples, batch 2,
    model/classifier in train mode as inside train() (pretrain_BreastPathQ.py:30-31,54-55,66)."""
    torch.manual_seed(42)
    model = quiet(refnet.TripletNet, "resnet18")
    cls = quiet(refnet.Classifier, 768, 6)
    model.train(); cls.train()
    i1, i2, i3 = (O.synthetic_patches(8, 224, seed=s) for s in (0, 1, 2))
    feats, logits = [], []
    with torch.no_grad():
        for b in range(4):
            sl = slice(2 * b, 2 * b + 2)
            f = model(i1[sl], i2[sl], i3[sl])
            feats.append(f)
            logits.append(cls(f))
    feats, logits = torch.cat(feats), torch.cat(logits)
    np.savez(os.path.join(OUT, "cfg1_rsp_forward.npz"), feats=np_(feats), logits=np_(logits),
             pred=np_(torch.argmax(logits, 1)), buffers=buffers_fp([model]))


def golden_rsp_step():
    """One full RSP train step (pretrain_BreastPathQ.py:53-68) at N=2 triples, 64x64."""
    torch.manual_seed(42)
    model = quiet(refnet.TripletNet, "resnet18")
    cls = quiet(refnet.Classifier, 768, 6)
    model.train(); cls.train()
    opt = torch.optim.SGD(list(model.parameters()) + list(cls.parameters()), lr=0.01, momentum=0.9,
                          weight_decay=1e-4, nesterov=True)
    i1, i2, i3 = (O.synthetic_patches(2, 64, seed=s) for s in (0, 1, 2))
    target = torch.tensor([3, 5])
    feats = model(i1, i2, i3)
    output = cls(feats)
    loss = torch.nn.CrossEntropyLoss()(output, target)
    opt.zero_grad(); loss.backward(); opt.step()
    np.savez(os.path.join(OUT, "rsp_step.npz"), loss=np_(loss), output=np_(output),
             pred=np_(torch.argmax(output, 1)), feats=np_(feats),
             grad_norms=grad_norms([model, cls]), params_after=params_fp([model, cls]),
             buffers=buffers_fp([model]))


def _cr_models(num_classes, seed=42):
    torch.manual_seed(seed)
    student = quiet(refnet.TripletNet_Finetune, "resnet18")
    cls_s = quiet(refnet.FinetuneResNet, num_classes)
    import copy
    teacher, cls_t = copy.deepcopy(student), copy.deepcopy(cls_s)  # both load the same ckpt (:394-402)
    for p in list(teacher.parameters()) + list(cls_t.parameters()):
        p.requires_grad = False
    teacher.eval(); cls_t.eval(); student.train(); cls_s.train()
    return teacher, student, cls_t, cls_s


def golden_cr_step(kind):
    """One consistency step (eval_BreastPathQ_SSL_CR.py:76-105 / eval_Kather_SSL_CR.py:71-105),
    b=1 labeled item (3 views), mu=2, 64x64, --modules_student 0 (full backward)."""
    C = 1 if kind == "mse" else 9
    teacher, student, cls_t, cls_s = _cr_models(C)
    opt = torch.optim.Adam([p for p in list(student.parameters()) + list(cls_s.parameters())
                            if p.requires_grad], lr=1e-4, betas=(0.9, 0.999), weight_decay=1e-4)
    inputs_x = O.synthetic_patches(3, 64, seed=10)
    inputs_u_w = O.synthetic_patches(2, 64, seed=11)
    inputs_u_s = O.synthetic_patches(2, 64, seed=12)
    if kind == "mse":
        targets_x = torch.tensor([0.25, 0.25, 0.25])
    else:
        targets_x = torch.tensor([4, 4, 4])
    with torch.no_grad():
        logits_u_w = cls_t(teacher(inputs_u_w))
    logits = cls_s(student(torch.cat((inputs_x, inputs_u_s))))
    lx, lus = logits[:3], logits[3:]
    if kind == "mse":
        sup = F.mse_loss(lx, targets_x.view(-1, 1), reduction="mean")
        cons = F.mse_loss(logits_u_w, lus, reduction="mean")
        extra = {}
    else:
        sup = F.cross_entropy(lx, targets_x, reduction="mean")
        pseudo = torch.softmax(logits_u_w.detach_(), dim=-1)
        _, targets_u = torch.max(pseudo, dim=-1)
        cons = F.cross_entropy(lus, targets_u, reduction="mean")
        extra = {"targets_u": np_(targets_u)}
    final = sup + 1.0 * cons
    opt.zero_grad(); final.backward(); opt.step()
    np.savez(os.path.join(OUT, "cr_step_%s.npz" % kind), sup=np_(sup), cons=np_(cons),
             final=np_(final), logits_x=np_(lx), logits_u_s=np_(lus), logits_u_w=np_(logits_u_w),
             grad_norms=grad_norms([student, cls_s]), params_after=params_fp([student, cls_s]),
             buffers=buffers_fp([student]), **extra)


def golden_finetune_step():
    """eval_Kather_SSL.py:62-79, 9 classes, N=4, 64x64, Adam(lr 1e-5, wd 1e-4)."""
    torch.manual_seed(42)
    model = quiet(refnet.TripletNet_Finetune, "resnet18")
    cls = quiet(refnet.FinetuneResNet, 9)
    model.train(); cls.train()
    opt = torch.optim.Adam(list(model.parameters()) + list(cls.parameters()), lr=1e-5,
                           betas=(0.9, 0.999), weight_decay=1e-4)
    x = O.synthetic_patches(4, 64, seed=20)
    target = torch.tensor([0, 8, 3, 3])
    output = cls(model(x))
    loss = torch.nn.CrossEntropyLoss()(output, target)
    opt.zero_grad(); loss.backward(); opt.step()
    np.savez(os.path.join(OUT, "finetune_step.npz"), loss=np_(loss), output=np_(output),
             pred=np_(torch.argmax(output, 1)), grad_norms=grad_norms([model, cls]),
             buffers=buffers_fp([model]),
             nbt=np.array([int(model.model.bn1.num_batches_tracked)]))


def golden_lookahead():
    """The reference's Lookahead wrapper (lookahead.py:81-106) around SGD on three small tensors,
    stepped 7 times with constant gradients: the pull happens on calls 5 (la_steps=5)."""
    g = torch.Generator().manual_seed(5)
    params = [torch.nn.Parameter(torch.randn(s, generator=g)) for s in ((7,), (3, 5), (2, 2, 2))]
    grads = [torch.randn(p.shape, generator=g) for p in params]
    init = torch.cat([p.detach().flatten() for p in params]).clone()
    inner = torch.optim.SGD(params, lr=0.1)
    import warnings
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        la = Lookahead(inner, la_steps=5, la_alpha=0.5)
        trace = []
        for _ in range(7):
            for p, gr in zip(params, grads):
                p.grad = gr.clone()
            la.step()
            trace.append(torch.cat([p.detach().flatten() for p in params]).clone())
    np.savez(os.path.join(OUT, "lookahead.npz"), trace=np_(torch.stack(trace)),
             init=np_(init),
             grads=np_(torch.cat([gr.flatten() for gr in grads])))


if __name__ == "__main__":
    os.makedirs(OUT, exist_ok=True)
    torch.set_num_threads(8)
    golden_init()
    golden_cfg1()
    golden_rsp_step()
    golden_cr_step("mse")
    golden_cr_step("ce")
    golden_finetune_step()
    golden_lookahead()
    for f in sorted(os.listdir(OUT)):
        print(f, os.path.getsize(os.path.join(OUT, f)))
