"""CPU tests of the host-side logic: C-ABI surface, module contract (names / order / init /
freezing / deepcopy / error behaviour), no-CPU-fallback guarantee, gradient arena + gloo
all-reduce with world_size 2."""
import copy
import ctypes
import os
import re
import socket
import sys

import numpy as np
import pytest
import torch
import torch.multiprocessing as mp

import ssl_cr_histo_b200 as b2n
from ssl_cr_histo_b200 import _lib, build, ddp, net
from oracle import ref_net as O
from util import golden

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


# ------------------------------------------------------------------ C ABI
def test_library_builds_loads_and_exports_every_declared_symbol():
    path = build.build_lib()
    lib = ctypes.CDLL(path)
    header = open(os.path.join(ROOT, "include", "b2n.h")).read()
    declared = sorted(set(re.findall(r"\b(b2n_[a-z0-9_]+)\s*\(", header)))
    assert len(declared) >= 25
    for name in declared:
        assert hasattr(lib, name), "libb2n.so does not export %s" % name
    assert sorted(_lib.EXPORTS) == declared, "python binding and header disagree"
    lib.b2n_version.restype = ctypes.c_int
    assert lib.b2n_version() == 2


def test_header_is_plain_c():
    """include/b2n.h is the drop-in boundary: it must compile as C (no C++ or CUDA types)."""
    import shutil
    import subprocess
    import tempfile

    if shutil.which("gcc") is None:
        pytest.skip("gcc not available")
    with tempfile.NamedTemporaryFile("w", suffix=".c") as f:
        f.write('#include "b2n.h"\nint (*probe)(void) = b2n_version;\nint main(void) { return probe == 0; }\n')
        f.flush()
        r = subprocess.run(["gcc", "-std=c99", "-Wall", "-Werror", "-pedantic", "-fsyntax-only",
                            "-I", os.path.join(ROOT, "include"), f.name], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr


def test_library_has_blackwell_tensor_core_and_tma_sass():
    import shutil
    import subprocess

    if shutil.which("cuobjdump") is None:
        pytest.skip("cuobjdump not available")
    sass = subprocess.run(["cuobjdump", "-sass", build.build_lib()], capture_output=True,
                          text=True).stdout
    assert "UTCHMMA" in sass or "UTCMMA" in sass or re.search(r"UTC\w*MMA", sass)  # tcgen05.mma
    assert "UTMALDG" in sass and "IM2COL" in sass                                    # TMA im2col
    assert "LDTM" in sass                                                            # tcgen05.ld
    assert "sm_100a" in subprocess.run(["cuobjdump", "-lelf", build.build_lib()],
                                       capture_output=True, text=True).stdout


# ------------------------------------------------------------- module contract
def test_constructor_signatures_and_errors_match_reference():
    with pytest.raises(NotImplementedError, match="not supported model type: vgg"):
        net.TripletNet("vgg")
    with pytest.raises(NotImplementedError, match="not supported model type: resnet34"):
        net.TripletNet_Finetune("resnet34")
    c = net.Classifier(768, 6)
    assert [tuple(p.shape) for p in c.parameters()] == [(128, 768), (128,), (6, 128), (6,)]
    f = net.FinetuneResNet(9)
    assert [tuple(p.shape) for p in f.parameters()] == [(9, 768), (9,)]
    assert list(f.state_dict().keys()) == ["classifier.0.weight", "classifier.0.bias"]


def test_parameter_names_order_and_seeded_init_equal_the_reference():
    g = golden("init_seed42.npz")
    torch.manual_seed(42)
    m = net.TripletNet("resnet18")
    c = net.Classifier(768, 6)
    assert [n for n, _ in m.named_parameters()] == list(g["param_names"])
    assert list(m.state_dict().keys()) == list(g["keys"])
    assert list(c.state_dict().keys()) == list(g["keys_cls"])
    for row, (k, v) in zip(list(g["fp"]) + list(g["fp_cls"]),
                           list(m.state_dict().items()) + list(c.state_dict().items())):
        f = v.double().flatten()
        assert abs(float(f.sum()) - row[0]) <= 1e-9 * max(1.0, abs(row[0])), k
        assert np.array_equal(f[:4].numpy(), row[2:2 + min(4, f.numel())]), k
    idx = {n: i for i, (n, _) in enumerate(m.named_parameters())}
    # the indices the reference's --modules flags freeze against (eval_Kather_SSL.py:229)
    assert idx["model.layer1.0.conv1.weight"] == 3 and idx["model.layer2.0.conv1.weight"] == 15
    assert idx["model.layer3.0.conv1.weight"] == 30 and idx["model.layer4.0.conv1.weight"] == 45
    assert idx["fc.0.weight"] == 60 and len(idx) == 64
    assert len(list(m.buffers())) == 60


def test_state_dict_round_trip_with_oracle_and_module_prefix_strip():
    ref = O.TripletNet_Finetune("resnet18")
    mine = net.TripletNet_Finetune("resnet18")
    mine.load_state_dict(ref.state_dict())                      # strict
    prefixed = {"module." + k: v for k, v in mine.state_dict().items()}
    stripped = {k[7:]: v for k, v in prefixed.items()}          # eval_Kather_SSL.py:344-346
    ref.load_state_dict(stripped)


def test_index_based_freezing_and_deepcopy():
    student = net.TripletNet_Finetune("resnet18")
    O.freeze_by_index(student, 60)                              # --modules_student 60 default
    flags = [p.requires_grad for p in student.parameters()]
    assert flags == [False] * 60 + [True] * 4
    teacher = copy.deepcopy(student)
    assert [p.requires_grad for p in teacher.parameters()] == flags
    for (k1, v1), (k2, v2) in zip(teacher.state_dict().items(), student.state_dict().items()):
        assert k1 == k2 and torch.equal(v1, v2) and v1.data_ptr() != v2.data_ptr()
    assert teacher.model._packs is not student.model._packs and not teacher.model._packs.entries
    assert "_packs" not in "".join(student.state_dict().keys())


def test_pickled_module_drops_the_packed_weight_cache():
    import io
    m = net.TripletNet_Finetune("resnet18")
    m.model._packs.entries["x"] = (None, torch.zeros(3), 1)
    buf = io.BytesIO()
    torch.save(m, buf)
    buf.seek(0)
    m2 = torch.load(buf, weights_only=False)
    assert m2.model._packs.entries == {} and m2.model._packs.generation == 1
    assert all(torch.equal(a, b) for a, b in zip(m.state_dict().values(), m2.state_dict().values()))


def test_pack_cache_staleness_rules():
    """The cache the multi-tensor prefill and the lazy per-conv path share: an entry is fresh for the
    tensor state it was built from, stale after an in-place write autograd sees, after this
    package's optimizer / lerp epochs, and after invalidate() (p.data writes, graph capture)."""
    from ssl_cr_histo_b200 import _lib
    from ssl_cr_histo_b200.trunk import _PackCache
    cache, p = _PackCache(), torch.nn.Parameter(torch.ones(4))
    calls = []

    def maker(w):
        calls.append(1)
        return w.clone()

    assert cache.stale("k", p)
    a = cache.get("k", p, maker)
    assert not cache.stale("k", p) and cache.get("k", p, maker) is a and len(calls) == 1
    cache.put("other", p, "packed elsewhere")                    # what the multi-tensor prefill does
    assert cache.get("other", p, maker) == "packed elsewhere" and len(calls) == 1
    with torch.no_grad():
        p.add_(1.0)                                              # bumps p._version
    assert cache.stale("k", p) and cache.get("k", p, maker) is not a and len(calls) == 2
    p._b2n_epoch = 7                                             # fused optimizer kernels
    assert cache.stale("k", p)
    cache.get("k", p, maker)
    _lib.WEIGHT_EPOCH += 1                                       # weights.lerp_
    assert cache.stale("k", p)
    cache.get("k", p, maker)
    p.data.mul_(0.5)                                             # invisible to every tag ...
    assert not cache.stale("k", p)
    cache.invalidate()                                           # ... hence the explicit invalidation
    assert cache.stale("k", p) and cache.stale("other", p)
    b = cache.get("k", p, maker)
    assert torch.equal(b, p.detach()) and not cache.stale("k", p)


def test_no_cpu_fallback():
    m, c = net.TripletNet("resnet18"), net.Classifier(768, 6)
    x = torch.zeros(1, 3, 32, 32)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        m(x, x, x)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        c(torch.zeros(2, 768))
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        b2n.losses.cross_entropy(torch.zeros(2, 6), torch.zeros(2, dtype=torch.long))
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        b2n.weights.lerp_([torch.zeros(3)], [torch.ones(3)], 0.5)


def test_product_package_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "ssl_cr_histo_b200")
    for fn in os.listdir(pkg):
        if fn.endswith(".py"):
            src = open(os.path.join(pkg, fn)).read()
            assert "oracle" not in src.replace("no CPU fallback", ""), fn
            assert "torchvision" not in src.replace("torchvision's", "").replace(
                "torchvision.models", "tv.models") or fn in ("trunk.py", "net.py"), fn
    for fn in ("trunk.py", "net.py", "heads.py"):
        src = open(os.path.join(pkg, fn)).read()
        assert "import torchvision" not in src and "F.conv2d" not in src and "cudnn" not in src


# -------------------------------------------------------------- DDP (gloo, CPU)
def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _ddp_worker(rank, world, port, out):
    import torch.distributed as dist

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.manual_seed(rank)                                      # replicas start DIFFERENT ...
    lin, lin2 = torch.nn.Linear(5, 3), torch.nn.Linear(3, 2)
    frozen = torch.nn.Parameter(torch.ones(2), requires_grad=False)
    params = list(lin.parameters()) + list(lin2.parameters())
    # two buckets (32-byte limit), started as their gradients are reported ready
    red = ddp.GradAllReducer(params + [frozen], overlap=True, bucket_mb=32 / 2 ** 20)
    out["w%d" % rank] = lin.weight.detach().clone()              # ... and are synchronised to rank 0
    assert red.nbytes == (15 + 3 + 6 + 2) * 4 and len(red.buckets) == 2
    assert lin2.bias.grad.data_ptr() == red.flat.data_ptr()      # arena in backward order
    assert lin.weight._b2n_grad_slot.data_ptr() == lin.weight.grad.data_ptr()
    data = torch.arange(40, dtype=torch.float32).view(8, 5) / 10.0
    shard = ddp.shard(data, rank, world)
    for step in range(3):                                        # re-armed by zero_grad
        red.zero_grad()
        lin2(lin(shard)).pow(2).mean().backward()                # autograd accumulates into the slots
        for p in reversed(params):                               # what the package's backward reports:
            red.ready(p)                                         # lin2 once, lin twice per step
        for p in lin.parameters():
            red.ready(p)
        # step 0 only counts the writers per slot; from step 1 on both buckets are already in flight
        assert all(n == -1 for n in red._pending) == (step > 0)
        red.all_reduce()
    out[rank] = torch.cat([p.grad.flatten() for p in params]).clone()
    assert lin.weight.grad.data_ptr() == red._slots[id(lin.weight)].data_ptr()   # still in the arena
    # a plain (non-overlapped) reducer and set_to_none
    red2 = ddp.GradAllReducer(params, sync_initial=False)
    for p in params:
        p.grad = None
    red2.zero_grad()
    lin2(lin(shard)).pow(2).mean().backward()
    red2.all_reduce(average=False)
    out["sum%d" % rank] = torch.cat([p.grad.flatten() for p in params]).clone()
    dist.destroy_process_group()


def test_grad_arena_allreduce_world2_matches_full_batch():
    world, port = 2, _free_port()
    mgr = mp.get_context("spawn").Manager()
    out = mgr.dict()
    mp.spawn(_ddp_worker, args=(world, port, out), nprocs=world, join=True)
    torch.manual_seed(0)
    lin, lin2 = torch.nn.Linear(5, 3), torch.nn.Linear(3, 2)
    assert torch.equal(out["w0"], lin.weight.detach()) and torch.equal(out["w1"], out["w0"])
    data = torch.arange(40, dtype=torch.float32).view(8, 5) / 10.0
    lin2(lin(data)).pow(2).mean().backward()
    full = torch.cat([p.grad.flatten() for p in list(lin.parameters()) + list(lin2.parameters())])
    assert torch.allclose(out[0], out[1])
    assert torch.allclose(out[0], full, rtol=1e-5, atol=1e-6)    # mean of shard means == full mean
    assert torch.allclose(out["sum0"], 2 * full, rtol=1e-5, atol=1e-6)   # average=False leaves the sum


def test_shard_rejects_ragged_batches():
    with pytest.raises(RuntimeError, match="does not divide"):
        ddp.shard(torch.zeros(7, 2), 0, 2)
    assert ddp.shard(torch.arange(8), 1, 4).tolist() == [2, 3]


def test_fused_optimizers_keep_the_torch_optim_contract():
    """Constructor validation, param_groups / LR-scheduler / state_dict compatibility (no launch)."""
    import pytest
    import torch
    from ssl_cr_histo_b200 import optim as fused

    p = [torch.nn.Parameter(torch.zeros(4))]
    adam = fused.Adam(p, lr=1e-4, weight_decay=1e-4)                   # eval_BreastPathQ_SSL_CR.py:481
    ref_defaults = torch.optim.Adam(p, lr=1e-4, weight_decay=1e-4).defaults
    assert all(ref_defaults[k] == v for k, v in adam.defaults.items())   # same names, same values
    sched = torch.optim.lr_scheduler.MultiStepLR(adam, milestones=[30, 60], gamma=0.1)   # :482
    assert sched.get_last_lr() == [1e-4]
    sgd = fused.SGD(p, lr=0.01, momentum=0.9, nesterov=True, weight_decay=1e-4)   # pretrain_BreastPathQ.py:245
    assert sgd.param_groups[0]["nesterov"] is True
    assert set(adam.state_dict().keys()) == {"state", "param_groups"}
    with pytest.raises(ValueError):
        fused.Adam(p, lr=-1.0)
    with pytest.raises(ValueError):
        fused.SGD(p, lr=0.1, nesterov=True)                             # Nesterov needs momentum
    with pytest.raises(NotImplementedError):
        fused.SGD(p, lr=0.1, momentum=0.9, dampening=0.5)
    p[0].grad = torch.zeros(4)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        adam.step()                                                     # CPU tensors fail loudly


def test_python_binding_argument_counts_match_the_header():
    """Every ctypes signature in _lib.py has exactly the parameters include/b2n.h declares (the last
    one being `void* stream`) -- a drifted binding would silently shift arguments."""
    header = open(os.path.join(ROOT, "include", "b2n.h")).read()
    header = re.sub(r"/\*.*?\*/", "", header, flags=re.S)          # strip comments
    for name, sig in _lib._SIGNATURES.items():
        m = re.search(r"\bint\s+%s\s*\((.*?)\)\s*;" % name, header, flags=re.S)
        assert m, "no declaration of %s in b2n.h" % name
        params = [p.strip() for p in m.group(1).split(",")]
        assert params[-1].replace(" ", "") == "void*stream", name
        assert len(params) == len(sig) + 1, "%s: header has %d parameters, binding %d" % (
            name, len(params), len(sig) + 1)
        for decl, ct in zip(params, sig):                            # pointers vs scalars line up
            is_ptr = "*" in decl
            assert is_ptr == (ct is ctypes.c_void_p), "%s: `%s` bound as %s" % (name, decl, ct.__name__)


def test_bench_reference_arm_prints_the_contract_line():
    """`bench.py --impl reference` (the CPU arm the driver times next to the CUDA arm) exits 0 and
    prints one JSON line with the contract's keys; non-zero ranks stay silent."""
    import json
    import subprocess

    env = dict(os.environ, RANK="0", OMP_NUM_THREADS="4")
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference",
                          "--steps", "1", "--warmup", "0", "--size", "64"], capture_output=True,
                         text=True, env=env, timeout=600)
    assert out.returncode == 0, out.stderr[-500:]
    line = json.loads(out.stdout.strip().splitlines()[-1])
    for key in ("impl", "metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step",
                "higher_is_better", "scaling", "vs_baseline", "dtype", "data", "config", "cpu_baseline",
                "e2e"):
        assert key in line, key
    assert line["impl"] == "reference" and line["unit"] == "patches/s" and line["value"] > 0
    assert line["cpu_baseline"]["kind"] == "port" and line["cpu_baseline"]["cores"] >= 1
    assert line["e2e"]["h2d_bytes_per_step"] == 0
    silent = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference",
                             "--steps", "1", "--warmup", "0", "--size", "64"], capture_output=True,
                            text=True, env=dict(env, RANK="1"), timeout=600)
    assert silent.returncode == 0 and silent.stdout.strip() == ""
