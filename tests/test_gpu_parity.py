"""End-to-end GPU parity: the drop-in classes driven by the reference's own (restated) loop
bodies, against (a) golden vectors produced by the REAL reference modules and (b) the CPU oracle
on the same seeded inputs.

Tolerances (BASELINE.json north_star / SURVEY.md section 8d):
  * fp32 logits and losses: max|d| / max|ref| <= 1e-3 (measured ~1e-5: forward convs run
    error-compensated 3xTF32);
  * RSP argmax: bit-exact, with an asserted top-2 margin floor;
  * BN running statistics: <= 1e-3 relative (measured ~3e-5), counters exact;
  * gradients: <= 2.5e-2 relative L2 per tensor; the largest value any test here measures is 1.57e-2
    (gpurun_out/grad_worst.jsonl; 9.3e-3 at full cfg3 size, 1.6e-2 at full cfg2 size).  SURVEY 8(d)
    asked for 1e-3; tools/grad_gate.py (profiles/r2_grad_gate.md, N=8 at 224x224) shows that no
    implementation meets that against an fp32 oracle: the oracle in float64 -- the exact answer -- is
    itself 4.3e-3 (worst tensor) / 2.5e-3 (median) away from the fp32 oracle, torch + cuDNN in strict
    fp32 on the B200 6.4e-3 / 4.1e-3, this package 5.6e-3 / 2.9e-3, and torch's own GPU default (TF32
    convolutions) 1.3e-1 / 9.8e-2.  The cause is not the arithmetic of the backward pass (TF32
    operand rounding of the backward convs alone: 1.3e-3, oracle with TF32_BACKWARD) but ReLU gates:
    a forward difference of ~1e-6 already flips some, and a gradient of random-sign terms sees that
    as sqrt(flipped fraction).
    DESIGN.md, "Numerics".
"""
import copy

import numpy as np
import pytest
import torch
import torch.nn.functional as F

import ssl_cr_histo_b200.net as net
from ssl_cr_histo_b200 import weights
from oracle import ref_net as O
from util import golden, max_rel, rel_l2, top2_margin

pytestmark = pytest.mark.gpu
DEV = "cuda"
TOL = 1e-3
GRAD_TOL = 2.5e-2


def pair(kind, head, seed=42):
    """(oracle model, oracle head, CUDA model, CUDA head) with the reference's seeded init."""
    st, hs = O.reference_state(seed, head)
    om = O.TripletNet("resnet18") if kind == "triplet" else O.TripletNet_Finetune("resnet18")
    gm = net.TripletNet("resnet18") if kind == "triplet" else net.TripletNet_Finetune("resnet18")
    if head[0] == "classifier":
        oh, gh = O.Classifier(768, head[1]), net.Classifier(768, head[1])
    else:
        oh, gh = O.FinetuneResNet(head[1]), net.FinetuneResNet(head[1])
    for m in (om, gm):
        m.load_state_dict(st)
    for h in (oh, gh):
        h.load_state_dict(hs)
    return om, oh, gm.to(DEV), gh.to(DEV)


def grads_of(mods):
    return [(n, p.grad) for m in mods for n, p in m.named_parameters()]


def assert_grads_close(mine, ref, tol=GRAD_TOL):
    worst = ("", 0.0)
    for (n1, g1), (n2, g2) in zip(grads_of(mine), grads_of(ref)):
        assert n1 == n2
        assert (g1 is None) == (g2 is None), n1
        if g1 is None:
            continue
        r = rel_l2(g1, g2)
        if r > worst[1]:
            worst = (n1, r)
    try:        # measured floors of every gradient comparison, for DESIGN.md / the tolerance
        import inspect, json, os
        out = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "gpurun_out")
        os.makedirs(out, exist_ok=True)
        with open(os.path.join(out, "grad_worst.jsonl"), "a") as f:
            f.write(json.dumps({"test": inspect.stack()[1].function, "tensor": worst[0],
                                "rel_l2": worst[1]}) + "\n")
    except OSError:
        pass
    assert worst[1] <= tol, "gradient %s differs by %.3e relative L2" % worst
    return worst[1]


def assert_buffers_close(mine, ref, tol=TOL):
    for (k1, b1), (k2, b2) in zip(mine.named_buffers(), ref.named_buffers()):
        assert k1 == k2
        if b1.dtype == torch.long:
            assert int(b1) == int(b2), k1
        else:
            assert max_rel(b1, b2) <= tol, k1


# ------------------------------------------------------------------ golden (real reference)
def test_cfg1_rsp_forward_matches_reference_golden():
    """BASELINE.json configs[0]: 8 synthetic 224x224 triples, batch 2, logits + permutation id."""
    g = golden("cfg1_rsp_forward.npz")
    _, _, model, cls = pair("triplet", ("classifier", 6))
    model.train(); cls.train()
    i1, i2, i3 = (O.synthetic_patches(8, 224, seed=s).to(DEV) for s in (0, 1, 2))
    feats, logits = [], []
    with torch.no_grad():
        for b in range(4):
            sl = slice(2 * b, 2 * b + 2)
            f = model(i1[sl], i2[sl], i3[sl])
            feats.append(f); logits.append(cls(f))
    feats, logits = torch.cat(feats), torch.cat(logits)
    assert max_rel(logits, g["logits"]) <= TOL
    assert max_rel(feats, g["feats"]) <= TOL
    margin = top2_margin(g["logits"])
    err = float((logits.cpu() - torch.tensor(g["logits"])).abs().max())
    assert margin > 10 * err, "top-2 margin %.2e too small against error %.2e" % (margin, err)
    assert np.array_equal(torch.argmax(logits, 1).cpu().numpy(), g["pred"])      # bit-exact argmax
    buf = np.array([float(b.double().sum()) for b in model.buffers()])
    assert np.allclose(buf, g["buffers"], rtol=TOL, atol=1e-2)
    assert int(model.model.bn1.num_batches_tracked) == 12                         # 4 iters x 3 passes


def test_step_goldens_losses_and_logits():
    g = golden("rsp_step.npz")
    _, _, model, cls = pair("triplet", ("classifier", 6))
    model.train(); cls.train()
    opt = O.make_rsp_optimizer(list(model.parameters()) + list(cls.parameters()))
    i1, i2, i3 = (O.synthetic_patches(2, 64, seed=s).to(DEV) for s in (0, 1, 2))
    out = O.rsp_pretrain_step(model, cls, opt, i1, i2, i3, torch.tensor([3, 5], device=DEV))
    assert abs(float(out["loss"]) - float(g["loss"])) <= TOL * abs(float(g["loss"]))
    assert max_rel(out["output"], g["output"]) <= TOL
    assert np.array_equal(out["pred"].cpu().numpy(), g["pred"])
    gn = np.array([float(p.grad.double().norm()) for m in (model, cls) for p in m.parameters()])
    assert np.allclose(gn, g["grad_norms"], rtol=GRAD_TOL, atol=1e-7)
    for kind, C in (("mse", 1), ("ce", 9)):
        g = golden("cr_step_%s.npz" % kind)
        _, _, student, cls_s = pair("finetune", ("finetune", C))
        teacher, cls_t = copy.deepcopy(student), copy.deepcopy(cls_s)
        O.freeze_by_index(teacher, 64)
        for p in cls_t.parameters():
            p.requires_grad = False
        teacher.eval(); cls_t.eval(); student.train(); cls_s.train()
        opt = O.make_cr_optimizer(list(student.parameters()) + list(cls_s.parameters()))
        tx = torch.tensor([0.25] * 3) if kind == "mse" else torch.tensor([4, 4, 4])
        out = O.consistency_step(teacher, student, cls_t, cls_s, opt,
                                 O.synthetic_patches(3, 64, seed=10).to(DEV), tx.to(DEV),
                                 O.synthetic_patches(2, 64, seed=11).to(DEV),
                                 O.synthetic_patches(2, 64, seed=12).to(DEV), 1.0, kind)
        for k, gk in (("sup", "sup"), ("cons", "cons"), ("loss", "final")):
            assert abs(float(out[k]) - float(g[gk])) <= TOL * max(abs(float(g[gk])), 1e-3), (kind, k)
        for k in ("logits_x", "logits_u_s", "logits_u_w"):
            assert max_rel(out[k], g[k]) <= TOL, (kind, k)
    g = golden("finetune_step.npz")
    _, _, model, cls = pair("finetune", ("finetune", 9))
    model.train(); cls.train()
    opt = torch.optim.Adam(list(model.parameters()) + list(cls.parameters()), lr=1e-5,
                           weight_decay=1e-4)
    out = O.finetune_step(model, cls, opt, O.synthetic_patches(4, 64, seed=20).to(DEV),
                          torch.tensor([0, 8, 3, 3], device=DEV))
    assert abs(float(out["loss"]) - float(g["loss"])) <= TOL * abs(float(g["loss"]))
    assert max_rel(out["output"], g["output"]) <= TOL
    assert np.array_equal(out["pred"].cpu().numpy(), g["pred"])
    assert int(model.model.bn1.num_batches_tracked) == 3


# ------------------------------------------------------------------ oracle, same inputs
@pytest.mark.parametrize("N,size", [(8, 224), (6, 96), (3, 256)])
def test_rsp_pretrain_step_parity(N, size):
    """cfg2 shape class (pretrain_BreastPathQ.py:53-68): logits / loss / argmax / BN buffers /
    gradients / the SGD-Nesterov-updated weights.  256 is the reference's default tile size
    (pretrain_BreastPathQ.py:188-189)."""
    i1, i2, i3 = (O.synthetic_patches(N, size, seed=s) for s in (0, 1, 2))
    target = torch.randint(0, 6, (N,), generator=torch.Generator().manual_seed(5))
    om, oh, gm, gh = pair("triplet", ("classifier", 6))
    for m in (om, oh, gm, gh):
        m.train()
    opt_g = O.make_rsp_optimizer(list(gm.parameters()) + list(gh.parameters()))
    out_g = O.rsp_pretrain_step(gm, gh, opt_g, i1.to(DEV), i2.to(DEV), i3.to(DEV), target.to(DEV))
    opt_o = O.make_rsp_optimizer(list(om.parameters()) + list(oh.parameters()))
    out_o = O.rsp_pretrain_step(om, oh, opt_o, i1, i2, i3, target)
    assert max_rel(out_g["output"], out_o["output"]) <= TOL
    assert max_rel(out_g["feats"], out_o["feats"]) <= TOL
    assert abs(float(out_g["loss"]) - float(out_o["loss"])) <= TOL * float(out_o["loss"])
    err = float((out_g["output"].cpu() - out_o["output"]).abs().max())
    assert top2_margin(out_o["output"]) > 10 * err
    assert torch.equal(out_g["pred"].cpu(), out_o["pred"])                 # bit-exact argmax
    assert_buffers_close(gm, om)
    assert_grads_close([gm, gh], [om, oh])
    for (n, p), (_, q) in zip(gm.named_parameters(), om.named_parameters()):
        # weights after the step (biases start at 0, so the bound is absolute there)
        assert float((p.cpu() - q).abs().max()) <= TOL * float(q.abs().max()) + 1e-5, n


@pytest.mark.parametrize("kind,C", [("mse", 1), ("ce", 9)])
def test_consistency_step_parity(kind, C):
    """cfg3 shape class (eval_BreastPathQ_SSL_CR.py:76-105 / eval_Kather_SSL_CR.py:71-105),
    --modules_student 0, b=2 labeled items x 3 views, mu=4, 224x224."""
    b, mu, size = 2, 4, 224
    ix = O.synthetic_patches(3 * b, size, seed=10)
    iw, is_ = O.synthetic_patches(b * mu, size, seed=11), O.synthetic_patches(b * mu, size, seed=12)
    g = torch.Generator().manual_seed(6)
    tx = torch.rand(3 * b, generator=g) if kind == "mse" else torch.randint(0, C, (3 * b,), generator=g)

    def run(models, dev):
        student, cls_s = models
        teacher, cls_t = copy.deepcopy(student), copy.deepcopy(cls_s)      # :394-402 same ckpt
        O.freeze_by_index(teacher, 64)
        for p in cls_t.parameters():
            p.requires_grad = False
        teacher.eval(); cls_t.eval(); student.train(); cls_s.train()
        opt = O.make_cr_optimizer(list(student.parameters()) + list(cls_s.parameters()))
        return O.consistency_step(teacher, student, cls_t, cls_s, opt, ix.to(dev), tx.to(dev),
                                  iw.to(dev), is_.to(dev), 1.0, kind)

    om, oh, gm, gh = pair("finetune", ("finetune", C))
    out_g = run((gm, gh), DEV)
    out_o = run((om, oh), "cpu")
    for k in ("logits_u_w", "logits_x", "logits_u_s"):       # teacher (eval) and student (train)
        assert max_rel(out_g[k], out_o[k]) <= TOL, k
    for k in ("sup", "cons", "loss"):
        assert abs(float(out_g[k]) - float(out_o[k])) <= TOL * max(abs(float(out_o[k])), 1e-3), k
    assert_buffers_close(gm, om)                             # incl. num_batches_tracked == 3
    assert_grads_close([gm, gh], [om, oh])


def test_frozen_trunk_default_and_partial_freeze():
    """--modules_student 60 (reference default: no conv backward at all) and --modules 45
    (layer4 + heads trainable): frozen parameters get no gradient, the rest match."""
    x = O.synthetic_patches(4, 96, seed=30)
    tgt = torch.tensor([1, 0, 1, 1])
    for n_frozen in (60, 45):
        om, oh, gm, gh = pair("finetune", ("finetune", 2))
        for m in (om, gm):
            O.freeze_by_index(m, n_frozen)
            m.train()
        lo = F.cross_entropy(oh(om(x)), tgt)
        lo.backward()
        lg = F.cross_entropy(gh(gm(x.to(DEV))), tgt.to(DEV))
        lg.backward()
        assert abs(float(lg) - float(lo)) <= 1e-4 * float(lo)
        for i, ((n, p), (_, q)) in enumerate(zip(gm.named_parameters(), om.named_parameters())):
            if i < n_frozen:
                assert p.grad is None and q.grad is None, n
            else:
                assert rel_l2(p.grad, q.grad) <= GRAD_TOL, n
        assert_buffers_close(gm, om)                         # BN stays in train mode while frozen


def test_finetune_single_pass_equals_three_reference_passes():
    """models/net.py:86-90 runs the trunk 3x on the same input; the CUDA module runs it once.
    Feature blocks, summed gradients and the 3-fold BN buffer update must all agree."""
    x = O.synthetic_patches(4, 96, seed=40)
    om, oh, gm, gh = pair("finetune", ("finetune", 9))
    om.train(); gm.train()
    f = gm(x.to(DEV))
    assert torch.equal(f[:, :256], f[:, 256:512]) and torch.equal(f[:, :256], f[:, 512:])
    fo = om(x)
    assert max_rel(f, fo) <= 1e-4
    assert_buffers_close(gm, om, 1e-4)
    assert int(gm.model.layer3[1].bn2.num_batches_tracked) == 3
    tgt = torch.tensor([0, 3, 8, 8])
    F.cross_entropy(gh(f), tgt.to(DEV)).backward()
    F.cross_entropy(oh(fo), tgt).backward()
    assert_grads_close([gm, gh], [om, oh])


def test_triplet_permutation_invariant_and_eval_mode():
    """In eval mode f12 only depends on (i1, i2); equal inputs give equal blocks; results are
    deterministic; and the eval-mode (BN folded into the conv epilogue) features match."""
    om, _, gm, _ = pair("triplet", ("classifier", 6))
    gm.eval(); om.eval()
    a, b = O.synthetic_patches(2, 64, seed=1).to(DEV), O.synthetic_patches(2, 64, seed=2).to(DEV)
    with torch.no_grad():
        f_aab = gm(a, a, b)
        f_aaa = gm(a, a, a)
        f_again = gm(a, a, b)
        ref = om(a.cpu(), a.cpu(), b.cpu())
    assert torch.equal(f_aab, f_again)                                    # deterministic
    assert torch.equal(f_aaa[:, :256], f_aaa[:, 256:512]) and torch.equal(f_aaa[:, :256], f_aaa[:, 512:])
    assert torch.equal(f_aab[:, :256], f_aaa[:, :256])                    # f12 only depends on (i1,i2)
    assert torch.equal(f_aab[:, 256:512], f_aab[:, 512:])                 # f23 == f13 when i1 == i2
    assert max_rel(f_aab, ref) <= TOL


def test_eval_mode_with_trainable_trunk_is_rejected_loudly():
    _, _, gm, _ = pair("finetune", ("finetune", 2))
    gm.eval()
    with pytest.raises(NotImplementedError, match="eval-mode BatchNorm"):
        gm(O.synthetic_patches(1, 32, seed=3).to(DEV))
    gm.train()
    with pytest.raises(RuntimeError, match="even"):
        gm(torch.zeros(1, 3, 33, 32, device=DEV))


def test_teacher_handoff_and_weight_cache_invalidation():
    _, _, student, _ = pair("finetune", ("finetune", 1))
    teacher = copy.deepcopy(student).eval()
    x = O.synthetic_patches(2, 64, seed=50).to(DEV)
    with torch.no_grad():
        before = teacher(x)
        for p in student.parameters():
            p.mul_(1.01)
        student.model.bn1.running_mean.add_(0.5)
        student.model.bn1.num_batches_tracked.add_(7)
    weights.teacher_handoff_(teacher, student)                # alpha = 1, one launch
    for (k, v), (_, w) in zip(teacher.state_dict().items(), student.state_dict().items()):
        assert torch.equal(v, w), k                           # bit-exact, incl. int64 counters
    with torch.no_grad():
        after = teacher(x)
        expect = copy.deepcopy(student).eval()(x)             # the reference's deepcopy hand-off
    assert not torch.equal(before, after)                     # packed-weight cache was refreshed
    assert torch.equal(after, expect)


def test_large_batch_properties():
    """Size-independent checks at a BASELINE-scale batch (N=128 at 224x224, one trunk pass):
    batch rows are independent in eval mode (split == whole, bit-exact), and train-mode BN
    statistics / features of a duplicated batch equal those of the single batch."""
    _, _, gm, _ = pair("finetune", ("finetune", 9))
    x = O.synthetic_patches(128, 224, seed=60).to(DEV)
    gm.eval()
    with torch.no_grad():
        whole = gm(x)
        parts = torch.cat([gm(x[:48]), gm(x[48:])])
    assert torch.equal(whole, parts)
    assert torch.isfinite(whole).all()
    gm.train()
    a, b = copy.deepcopy(gm), copy.deepcopy(gm)
    with torch.no_grad():
        fa = a(x[:32])
        fb = b(torch.cat([x[:32], x[:32]]))
    assert max_rel(fb[:32], fa) <= 1e-4
    for (k, u), (_, v) in zip(a.named_buffers(), b.named_buffers()):
        if u.dtype != torch.long and "running_mean" in k:
            assert max_rel(u, v) <= 1e-4, k


def test_uint8_patches_equal_their_float_cast():
    """The trunk takes the patches as uint8 (what the dataset holds before the loop's .float(),
    dataset.py:65-67): features, gradients and BN buffers are bit-identical to the fp32 path --
    the uint8 values are exact in the stem's FP16 operand, whose lo plane is zero either way."""
    _, _, gm, gh = pair("finetune", ("finetune", 9))
    gm2, gh2 = copy.deepcopy(gm), copy.deepcopy(gh)
    xf = O.synthetic_patches(6, 96, seed=70)
    xu = xf.to(torch.uint8)
    assert torch.equal(xu.float(), xf)
    target = torch.tensor([0, 3, 8, 1, 5, 2], device=DEV)
    gm.train(); gm2.train()
    lf = F.cross_entropy(gh(gm(xf.to(DEV))), target)
    lu = F.cross_entropy(gh2(gm2(xu.to(DEV))), target)
    assert torch.equal(lf, lu)
    lf.backward(); lu.backward()
    for (k, p), (_, q) in zip(gm.named_parameters(), gm2.named_parameters()):
        assert torch.equal(q.grad, p.grad), k       # (deterministic split-K reduction is the default)
    for (k, u), (_, v) in zip(gm.named_buffers(), gm2.named_buffers()):
        assert torch.equal(u, v), k
    gm.eval(); gm2.eval()
    with torch.no_grad():
        assert torch.equal(gm(xf.to(DEV)), gm2(xu.to(DEV)))


def test_wsi_heatmap_inference_matches_reference_loop():
    """test_Camelyon16.py:30-70 restated on both sides: eval-mode TripletNet_Finetune +
    FinetuneResNet(2), softmax 'tumor' column scattered into the slide's probability map."""
    from ssl_cr_histo_b200 import infer
    om, oh, gm, gh = pair("finetune", ("finetune", 2))
    # a few train-mode passes first so the running statistics are not the 0 / 1 defaults
    warm = O.synthetic_patches(8, 64, seed=80)
    om.train(); gm.train()
    with torch.no_grad():
        om(warm); gm(warm.to(DEV))
    xs = [O.synthetic_patches(5, 64, seed=81 + i) for i in range(3)]
    coords = [(np.arange(5) + 5 * i, (np.arange(5) * 2 + i) % 7) for i in range(3)]
    om.eval(); oh.eval()
    ref_map = np.zeros((15, 7))
    with torch.no_grad():
        for x, (xm, ym) in zip(xs, coords):
            ref_map[xm, ym] = torch.softmax(oh(om(x)), dim=1)[:, -1].numpy()
    got = infer.probability_map(gm, gh, [(x, torch.from_numpy(xm), torch.from_numpy(ym))
                                         for x, (xm, ym) in zip(xs, coords)], (15, 7))
    assert not gm.training and not gh.training
    assert np.abs(got - ref_map).max() <= 1e-3, np.abs(got - ref_map).max()
    assert got[ref_map == 0].max() == 0
    u8 = infer.tumor_probabilities(gm, gh, xs[0].to(torch.uint8).to(DEV))
    assert torch.equal(u8, infer.tumor_probabilities(gm, gh, xs[0].to(DEV)))


@pytest.mark.parametrize("N,H,W", [(1, 64, 96), (3, 34, 46), (2, 32, 640)])
def test_single_patch_and_non_square_train_step(N, H, W):
    """Edge shapes through the whole path: a batch of one / odd pooled sizes / non-square patches /
    a patch too wide for the fused stem backward's shared-memory bands (unfused chain instead);
    full fine-tune step (forward, CE-9 loss, backward) against the oracle."""
    om, oh, gm, gh = pair("finetune", ("finetune", 9))
    g = torch.Generator().manual_seed(90 + N)
    x = torch.randint(0, 256, (N, 3, H, W), dtype=torch.uint8, generator=g).float()
    target = torch.randint(0, 9, (N,), generator=g)
    om.train(); gm.train()
    lo = F.cross_entropy(oh(om(x)), target)
    lg = F.cross_entropy(gh(gm(x.to(DEV))), target.to(DEV))
    assert abs(float(lg.detach()) - float(lo.detach())) <= TOL * abs(float(lo.detach()))
    lo.backward(); lg.backward()
    if N > 1:       # (with one sample and a 2x2 final map the BN statistics are too thin for a grad gate)
        assert_grads_close([gm, gh], [om, oh])
    assert_buffers_close(gm, om)


# ------------------------------------------------------------------ BASELINE batch sizes vs the oracle
def _need_host_memory(gb):
    import psutil
    avail = psutil.virtual_memory().available / 2 ** 30
    if avail < gb:
        pytest.skip("full-size CPU oracle needs ~%d GB of host memory, %.0f GB available" % (gb, avail))


def _report(name, rows):
    """Per-tensor gradient / buffer differences of a full-size run, kept next to the bench artefacts."""
    import json
    import os
    out = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "gpurun_out")
    os.makedirs(out, exist_ok=True)
    with open(os.path.join(out, "parity_%s.json" % name), "w") as f:
        json.dump(rows, f, indent=1)
    print("\n[%s] " % name + ", ".join("%s=%.3g" % (k, v) for k, v in rows["summary"].items()))


def _grad_table(mine, ref):
    return {n1: rel_l2(g1, g2) for (n1, g1), (_, g2) in zip(grads_of(mine), grads_of(ref)) if g1 is not None}


def test_cfg3_full_size_consistency_step_vs_oracle():
    """BASELINE.json configs[2] at its real size -- eval_BreastPathQ_SSL_CR.py:62-105 with
    --batch_size 64 --mu 8 --modules_student 0: 192 labeled + 512 weak + 512 strong 224x224
    patches, teacher eval / student train, MSE/MSE, Adam -- once against the CPU oracle (which runs
    the trunk three times per model call, models/net.py:88-90).  704-row tiles wrap the 148 SMs
    hundreds of times and the HALO rasters reach 2.4 M rows here; nothing smaller exercises that."""
    _need_host_memory(150)
    b, mu, size = 64, 8, 224
    ix = O.synthetic_patches(3 * b, size, seed=10)
    iw, is_ = O.synthetic_patches(b * mu, size, seed=11), O.synthetic_patches(b * mu, size, seed=12)
    tx = torch.rand(3 * b, generator=torch.Generator().manual_seed(6))

    def run(models, dev):
        student, cls_s = models
        teacher, cls_t = copy.deepcopy(student), copy.deepcopy(cls_s)
        O.freeze_by_index(teacher, 64)
        for p in cls_t.parameters():
            p.requires_grad = False
        teacher.eval(); cls_t.eval(); student.train(); cls_s.train()
        opt = O.make_cr_optimizer(list(student.parameters()) + list(cls_s.parameters()))
        before = [p.detach().clone() for p in list(student.parameters()) + list(cls_s.parameters())]
        out = O.consistency_step(teacher, student, cls_t, cls_s, opt, ix.to(dev), tx.to(dev),
                                 iw.to(dev), is_.to(dev), 1.0, "mse")
        return out, before

    om, oh, gm, gh = pair("finetune", ("finetune", 1))
    out_g, before_g = run((gm, gh), DEV)
    torch.cuda.synchronize()
    out_o, _ = run((om, oh), "cpu")
    summary = {}
    for k in ("logits_u_w", "logits_x", "logits_u_s"):
        summary[k] = max_rel(out_g[k], out_o[k])
        assert summary[k] <= TOL, (k, summary[k])
    for k in ("sup", "cons", "loss"):
        summary[k] = abs(float(out_g[k]) - float(out_o[k])) / max(abs(float(out_o[k])), 1e-3)
        assert summary[k] <= TOL, (k, summary[k])
    assert_buffers_close(gm, om)
    table = _grad_table([gm, gh], [om, oh])
    summary["grad_rel_l2_worst"] = max(table.values())
    summary["grad_rel_l2_median"] = sorted(table.values())[len(table) // 2]
    # Adam's first step moves every weight by ~lr * sign(g): the updated weights agree to a small
    # fraction of lr except where a near-zero gradient element has the other sign (bounded by 2 lr)
    lr, flips, total = 1e-4, 0, 0
    for (n, p), (_, q), p0 in zip(list(gm.named_parameters()) + list(gh.named_parameters()),
                                  list(om.named_parameters()) + list(oh.named_parameters()), before_g):
        d = (p.detach().cpu() - q.detach()).abs()
        assert float(d.max()) <= 2.1 * lr + 1e-6 * float(q.abs().max()), n
        flips += int((d > 0.25 * lr).sum()); total += d.numel()
        assert not torch.equal(p.detach(), p0), n           # the optimizer really stepped
    summary["adam_step_fraction_off_by_quarter_lr"] = flips / total
    assert flips / total < 0.02
    _report("cfg3_full", {"summary": summary, "grad_rel_l2": table})
    assert summary["grad_rel_l2_worst"] <= GRAD_TOL, max(table, key=table.get)


def test_cfg2_full_size_rsp_step_vs_oracle():
    """BASELINE.json configs[1] at its real size -- pretrain_BreastPathQ.py:42-68 with 256 triples
    (768 patches, three trunk passes with per-pass BN statistics), CE over the 6 orders,
    SGD-Nesterov: logits, loss, bit-exact argmax (margin-checked), BN buffers, gradients, updated
    weights against the CPU oracle."""
    _need_host_memory(64)
    N, size = 256, 224
    i1, i2, i3 = (O.synthetic_patches(N, size, seed=s) for s in (0, 1, 2))
    target = torch.randint(0, 6, (N,), generator=torch.Generator().manual_seed(5))
    om, oh, gm, gh = pair("triplet", ("classifier", 6))
    for m in (om, oh, gm, gh):
        m.train()
    opt_g = O.make_rsp_optimizer(list(gm.parameters()) + list(gh.parameters()))
    out_g = O.rsp_pretrain_step(gm, gh, opt_g, i1.to(DEV), i2.to(DEV), i3.to(DEV), target.to(DEV))
    torch.cuda.synchronize()
    opt_o = O.make_rsp_optimizer(list(om.parameters()) + list(oh.parameters()))
    out_o = O.rsp_pretrain_step(om, oh, opt_o, i1, i2, i3, target)
    summary = {"logits": max_rel(out_g["output"], out_o["output"]),
               "feats": max_rel(out_g["feats"], out_o["feats"]),
               "loss": abs(float(out_g["loss"]) - float(out_o["loss"])) / float(out_o["loss"])}
    assert summary["logits"] <= TOL and summary["feats"] <= TOL and summary["loss"] <= TOL, summary
    # bit-exact permutation id wherever the oracle's own top-2 margin exceeds 10x the logit error;
    # rows below that margin are ties at fp32 resolution and are counted, not compared
    err = float((out_g["output"].cpu() - out_o["output"]).abs().max())
    top = out_o["output"].double().topk(2, dim=1).values
    decided = (top[:, 0] - top[:, 1]) > 10 * err
    summary["argmax_rows_decided"] = float(decided.float().mean())
    assert summary["argmax_rows_decided"] >= 0.97
    assert torch.equal(out_g["pred"].cpu()[decided], out_o["pred"][decided])
    assert_buffers_close(gm, om)
    table = _grad_table([gm, gh], [om, oh])
    summary["grad_rel_l2_worst"] = max(table.values())
    summary["grad_rel_l2_median"] = sorted(table.values())[len(table) // 2]
    for (n, p), (_, q) in zip(gm.named_parameters(), om.named_parameters()):
        assert float((p.cpu() - q).abs().max()) <= TOL * float(q.abs().max()) + 1e-5, n
    _report("cfg2_full", {"summary": summary, "grad_rel_l2": table})
    assert summary["grad_rel_l2_worst"] <= GRAD_TOL, max(table, key=table.get)


def test_parameter_writes_through_data_are_seen_by_the_packed_weights():
    """The vendored Lookahead (lookahead.py:96-97) and RAdam write ``p.data``, which bumps no version
    counter.  A training-mode trunk repacks after every backward pass; an eval-mode one is told with
    ``invalidate_packs()``."""
    _, _, gm, gh = pair("finetune", ("finetune", 9))
    x = O.synthetic_patches(4, 64, seed=95).to(DEV)
    target = torch.tensor([0, 3, 8, 1], device=DEV)
    gm.train()
    F.cross_entropy(gh(gm(x)), target).backward()
    for p in gm.parameters():
        p.data.mul_(0.5)                                   # lookahead-style write, _version unchanged
    fresh = copy.deepcopy(gm)                              # packs are never copied: built from p.data
    with torch.no_grad():
        assert torch.equal(gm(x), fresh(x))
    gm.eval(); fresh.eval()
    with torch.no_grad():
        before = gm(x)
        for p in gm.parameters():
            p.data.mul_(1.5)
        gm.model.invalidate_packs()
        after = gm(x)
        expect = copy.deepcopy(gm).eval()(x)
    assert not torch.equal(before, after) and torch.equal(after, expect)


@pytest.mark.parametrize("kind", ["finetune", "triplet"])
def test_gradient_arena_slots_equal_autograd_accumulation(kind):
    """ddp.GradAllReducer on one rank is just the flat arena: the backward kernels accumulate every
    parameter gradient straight into its slot (three writers per slot for TripletNet's three trunk
    passes) and autograd sees None -- same values as the per-parameter tensors autograd would have
    summed."""
    from ssl_cr_histo_b200 import ddp
    head = ("finetune", 9) if kind == "finetune" else ("classifier", 6)
    _, _, gm, gh = pair(kind, head)
    gm2, gh2 = copy.deepcopy(gm), copy.deepcopy(gh)
    xs = [O.synthetic_patches(4, 64, seed=96 + i).to(DEV) for i in range(3 if kind == "triplet" else 1)]
    target = torch.tensor([0, 3, 5, 1], device=DEV)
    gm.train(); gm2.train()
    F.cross_entropy(gh(gm(*xs)), target).backward()
    params2 = list(gm2.parameters()) + list(gh2.parameters())
    red = ddp.GradAllReducer(params2)
    for _ in range(2):                                     # slots are re-zeroed, not re-allocated
        red.zero_grad()
        F.cross_entropy(gh2(gm2(*xs)), target).backward()
        red.all_reduce()
    for (n, p), q in zip(list(gm.named_parameters()) + list(gh.named_parameters()), params2):
        assert q.grad.data_ptr() == q._b2n_grad_slot.data_ptr(), n
        assert rel_l2(q.grad, p.grad) < 1e-5, n            # (three writers per slot: another summation order)


def test_cuda_graph_replay_equals_eager_steps():
    """graph.GraphedStep: a whole consistency step (teacher forward, student forward, fused loss,
    backward, multi-tensor optimizer step) captured once and replayed on new batches -- weights, BN
    buffers and losses follow the eager run of the same steps.  (SGD: linear in the gradient, so the
    split-K summation-order noise of the weight gradients stays at round-off; Adam's sign-like first
    steps would amplify it.  The capturable Adam is checked in test_gpu_kernels.py.)"""
    from ssl_cr_histo_b200 import graph, losses, optim

    def build():
        _, _, student, cls_s = pair("finetune", ("finetune", 1))
        teacher, cls_t = copy.deepcopy(student).eval(), copy.deepcopy(cls_s).eval()
        for p in list(teacher.parameters()) + list(cls_t.parameters()):
            p.requires_grad = False
        student.train(); cls_s.train()
        params = list(student.parameters()) + list(cls_s.parameters())
        opt = optim.SGD(params, lr=1e-5, momentum=0.9, nesterov=True, weight_decay=1e-4)

        def step(ix, tx, iw, is_):
            with torch.no_grad():
                lw = cls_t(teacher(iw))
            logits = cls_s(student(torch.cat((ix, is_))))
            loss, _ = losses.consistency_mse(logits[:ix.shape[0]], tx, lw, logits[ix.shape[0]:], 1.0)
            opt.zero_grad(set_to_none=True)
            loss.backward()
            opt.step()
            return loss
        return student, cls_s, step

    batches = [(O.synthetic_patches(3, 64, seed=200 + i).to(DEV), torch.rand(3, device=DEV),
                O.synthetic_patches(4, 64, seed=300 + i).to(DEV),
                O.synthetic_patches(4, 64, seed=400 + i).to(DEV)) for i in range(5)]
    s_g, c_g, step_g = build()
    g = graph.GraphedStep(step_g, batches[0], warmup=2)        # 2 real warm-up steps; capture runs nothing
    graph_losses = [float(g(*b)) for b in batches]
    # the graphed model took 2 extra steps on batch 0 first: compare a fresh eager model that did too
    s_r, c_r, step_r = build()
    for _ in range(2):
        step_r(*batches[0])
    ref_losses = [float(step_r(*b)) for b in batches]
    assert len(set(graph_losses)) == len(graph_losses)          # every replay saw its own batch
    for a, b in zip(graph_losses, ref_losses):
        assert abs(a - b) <= 1e-4 * max(abs(b), 1e-3), (graph_losses, ref_losses)
    # zero-initialised parameters (biases) are pure sums of lr * gradient: they inherit the
    # gradients' sensitivity to summation-order noise in the previous step (ReLU gates, see the
    # module docstring; 2e-3 measured), everything else agrees to round-off
    for (n, p), (_, q) in zip(s_g.named_parameters(), s_r.named_parameters()):
        assert rel_l2(p, q) < (2e-2 if n.endswith("bias") else 1e-4), n
    for (k, u), (_, v) in zip(s_g.named_buffers(), s_r.named_buffers()):
        assert (int(u) == int(v)) if u.dtype == torch.long else max_rel(u, v) < 1e-4, k


def test_deterministic_weight_gradients_are_bit_repeatable():
    """Default mode: split-K partial planes summed in a fixed order -- two runs of the same step give
    bit-identical gradients for every parameter; the atomic-reduction mode
    (trunk.set_deterministic(False)) agrees with it to round-off."""
    from ssl_cr_histo_b200 import trunk
    _, _, gm, gh = pair("finetune", ("finetune", 9))
    x = O.synthetic_patches(6, 96, seed=110).to(DEV)
    target = torch.tensor([0, 3, 8, 1, 5, 2], device=DEV)

    def grads(model, head):
        model, head = copy.deepcopy(model).train(), copy.deepcopy(head)
        F.cross_entropy(head(model(x)), target).backward()
        return [p.grad.clone() for p in list(model.parameters()) + list(head.parameters())]

    assert trunk.DETERMINISTIC_WGRAD
    a, b = grads(gm, gh), grads(gm, gh)
    trunk.set_deterministic(False)
    try:
        base = grads(gm, gh)
    finally:
        trunk.set_deterministic(True)
    for (n, _), u, v, w in zip(list(gm.named_parameters()) + list(gh.named_parameters()), a, b, base):
        assert torch.equal(u, v), n
        assert rel_l2(u, w) < 1e-5, n


def test_example_loop_runs_the_whole_pipeline():
    """examples/ssl_cr_synthetic.py: GPU augmentation -> teacher / student -> fused loss -> backward ->
    multi-tensor Adam -> per-epoch teacher hand-off, two epochs; losses finite and moving, the
    teacher equals the student after the hand-off."""
    import importlib.util
    import os
    path = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "examples", "ssl_cr_synthetic.py")
    spec = importlib.util.spec_from_file_location("ssl_cr_synthetic", path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    history, student, teacher = mod.main(epochs=2, steps_per_epoch=2, batch_size=2, mu=2, image_size=64,
                                         log=lambda *_: None)
    assert len(history) == 4 and all(np.isfinite(h).all() for h in history)
    assert history[0][0] != history[-1][0]
    for (k, v), (_, w) in zip(teacher.state_dict().items(), student.state_dict().items()):
        assert torch.equal(v, w), k


def test_side_stream_weight_gradients_are_bit_identical_to_the_single_stream_pass():
    """trunk.OVERLAP_WGRAD: 2 (default) runs every weight gradient on a side stream beside the next
    element-wise phase of the backward pass, 1 as soon as its operands exist, 0 on the main stream.
    Same kernels, same operands: the gradients must agree bit for bit, step after step."""
    from ssl_cr_histo_b200 import trunk
    _, _, gm, gh = pair("finetune", ("finetune", 9))
    x = O.synthetic_patches(6, 96, seed=121).to(DEV)
    target = torch.tensor([0, 3, 8, 1, 5, 2], device=DEV)

    def grads(mode):
        old, trunk.OVERLAP_WGRAD = trunk.OVERLAP_WGRAD, mode
        try:
            m, c = copy.deepcopy(gm).train(), copy.deepcopy(gh)
            out = []
            for _ in range(2):
                m.zero_grad(set_to_none=True); c.zero_grad(set_to_none=True)
                F.cross_entropy(c(m(x)), target).backward()
                torch.cuda.synchronize()
                out.append([p.grad.clone() for p in m.parameters()])
            return out
        finally:
            trunk.OVERLAP_WGRAD = old

    assert trunk.OVERLAP_WGRAD == 2
    ref = grads(0)
    for mode in (2, 1):
        got = grads(mode)
        for step, (a, b) in enumerate(zip(ref, got)):
            for (n, _), u, v in zip(gm.named_parameters(), a, b):
                assert torch.equal(u, v), (mode, step, n)


def test_merged_stride2_data_gradient_equals_the_per_class_launches():
    """The three stride-2 blocks' data gradients: one merged launch (four accumulators per tile)
    against the four per-class launches -- same taps in the same order, so the whole backward pass
    agrees bit for bit; and the default, which also accumulates the 1x1 shortcut's gradient inside
    the merged launch, against both."""
    from ssl_cr_histo_b200 import trunk
    _, _, gm, gh = pair("finetune", ("finetune", 9))
    x = O.synthetic_patches(5, 96, seed=120).to(DEV)
    target = torch.tensor([0, 3, 8, 1, 5], device=DEV)

    def grads():
        m, c = copy.deepcopy(gm).train(), copy.deepcopy(gh)
        F.cross_entropy(c(m(x)), target).backward()
        return [p.grad.clone() for p in m.parameters()]

    assert trunk.MERGED_S2_DGRAD and trunk.FUSED_S2_SHORTCUT
    fused = grads()                      # default: the 1x1 shortcut gradient accumulated in the merged launch
    trunk.FUSED_S2_SHORTCUT = False
    try:
        merged = grads()                 # merged launch + a 1x1 launch whose result it adds
        trunk.MERGED_S2_DGRAD = False
        per_class = grads()
    finally:
        trunk.MERGED_S2_DGRAD = True
        trunk.FUSED_S2_SHORTCUT = True
    for (n, _), a, b, c in zip(gm.named_parameters(), merged, per_class, fused):
        assert torch.equal(a, b), n
        # The fused shortcut only moves one fp32 addition (the shortcut products join the accumulator
        # instead of being added to its rounded result): dx differs by 2-7e-7 relative
        # (tools/dbg_s2sc.py).  Every BatchNorm-backward output below is re-rounded to TF32, and a
        # perturbation d flips a fraction d / 2^-10 of those roundings by a whole TF32 ulp, so the
        # difference grows by ~sqrt per layer until it sits at the TF32 rounding noise that the
        # backward pass carries anyway (measured: 1.5e-5 one block below, 9e-4 at the stem).
        assert float((c - a).norm() / a.norm().clamp_min(1e-30)) < 3e-3, n
