#!/usr/bin/env python
"""Headline benchmark: 224x224 histo patches/sec of the SSL_CR teacher-student consistency step.

    python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path
    python bench.py --impl reference --steps K --warmup W     # the reference's CPU path (oracle port)
    torchrun --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...   (N > 1, one rank per GPU)

Workload (BASELINE.json configs[2], SURVEY.md section 8d "cfg3"; per rank, weak scaling -- 8 ranks
are configs[3]): eval_BreastPathQ_SSL_CR.py:76-100 with --batch_size 64 --mu 8 --lambda_u 1
--modules_student 0: labeled inputs_x (192,3,224,224), weak / strong unlabeled (512,3,224,224)
each; frozen teacher (eval, no_grad) + FinetuneResNet(1) on the weak view, student (train) on
cat(labeled, strong), MSE/MSE consistency loss, full backward, Adam(lr 1e-4, wd 1e-4).
One step = 704 unique patches per rank.  Synthetic uint8-valued patches, seeded random-init
weights (the reference's constructors under torch.manual_seed(42)).

Prints ONE JSON line (rank 0).  `value`: inputs resident in HBM; `e2e`: same step through the
public modules with per-step pinned-host -> device input copies and a loss read-back.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

FLOP_FWD, FLOP_BWD = 3.627e9, 7.018e9   # per patch per trunk pass at 224^2


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=8)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--batch-size", type=int, default=64, help="labeled items per rank (x3 views)")
    ap.add_argument("--mu", type=int, default=8)
    ap.add_argument("--size", type=int, default=224)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-profile", action="store_true")
    ap.add_argument("--workload", default="cr", choices=["cr", "rsp"],
                    help="cr: SSL_CR consistency step (BASELINE configs[2], the headline metric); "
                         "rsp: RSP pretext step, --rsp-batch triples (BASELINE configs[1])")
    ap.add_argument("--rsp-batch", type=int, default=256, help="triples per rank for --workload rsp")
    ap.add_argument("--torch-optim", action="store_true",
                    help="step torch.optim.Adam instead of the multi-tensor ssl_cr_histo_b200.optim.Adam")
    ap.add_argument("--e2e-fp32", action="store_true",
                    help="e2e arm ships fp32 patches (the reference loop's format) instead of uint8")
    return ap.parse_args()


# ------------------------------------------------------------------ reference / CPU arm
def cpu_reference_step_rate(steps, warmup, b=2, mu=7, size=224):
    """The reference's consistency step on the host cores: the oracle port of models/net.py
    driven by the restated loop body (eval_BreastPathQ_SSL_CR.py:76-100).  Bounded sample:
    b labeled items x 3 views + b*mu weak + b*mu strong patches."""
    from oracle import ref_net as O

    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    torch.manual_seed(42)
    student, cls_s = O.TripletNet_Finetune("resnet18"), O.FinetuneResNet(1)
    teacher, cls_t = O.teacher_handoff(student), O.teacher_handoff(cls_s)
    O.freeze_by_index(teacher, 64)
    for p in cls_t.parameters():
        p.requires_grad = False
    teacher.eval(); cls_t.eval(); student.train(); cls_s.train()
    opt = O.make_cr_optimizer(list(student.parameters()) + list(cls_s.parameters()))
    ix = O.synthetic_patches(3 * b, size, seed=0)
    iw, is_ = O.synthetic_patches(b * mu, size, seed=1), O.synthetic_patches(b * mu, size, seed=2)
    tx = torch.rand(3 * b, generator=torch.Generator().manual_seed(3))
    times = []
    for i in range(warmup + steps):
        t0 = time.perf_counter()
        out = O.consistency_step(teacher, student, cls_t, cls_s, opt, ix, tx, iw, is_, 1.0, "mse")
        float(out["loss"])
        if i >= warmup:
            times.append(time.perf_counter() - t0)
    total = sum(times)
    patches = 3 * b + b * mu
    return {"value": patches * len(times) / total, "ms_per_step": 1e3 * total / len(times),
            "cores": cores, "patches_per_step": patches,
            "sample": "b=%d labeled items x3 views + %d weak + %d strong %dx%d patches per step, "
                      "%d timed steps after %d warm-up" % (b, b * mu, b * mu, size, size, len(times), warmup)}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    r = cpu_reference_step_rate(args.steps, args.warmup, size=args.size)
    line = {
        "impl": "reference", "metric": "224x224 histo patches/sec (consistency step)",
        "value": r["value"], "unit": "patches/s", "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": r["ms_per_step"], "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": "SSL_CR consistency step (eval_BreastPathQ_SSL_CR.py), MSE/MSE, "
                               "modules_student=0, CPU sample of cfg3", "sample": r["sample"]},
        "cpu_baseline": {"value": r["value"], "unit": "patches/s", "cores": r["cores"],
                         "kind": "port", "sample": r["sample"]},
        "e2e": {"value": r["value"], "unit": "patches/s", "h2d_bytes_per_step": 0,
                "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


# ------------------------------------------------------------------------- CUDA arm
class ClockSampler(threading.Thread):
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.rows, self.stop_flag = index, [], False

    def run(self):
        while not self.stop_flag:
            try:
                out = subprocess.run(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                      "--format=csv,noheader,nounits"], capture_output=True,
                                     text=True, timeout=5).stdout.strip()
                if out:
                    self.rows.append([c.strip() for c in out.split(",")])
            except Exception:
                pass
            time.sleep(0.02)

    def summary(self):
        sm = sorted(int(r[0]) for r in self.rows if r and r[0].isdigit())
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names)
                   if any(len(r) > 3 + i and r[3 + i].lower().startswith("active") for r in self.rows)]
        mx = [int(r[1]) for r in self.rows if len(r) > 1 and r[1].isdigit()]
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": reasons, "samples": len(self.rows)}


def run_b200(args):
    import torch.distributed as dist
    import ssl_cr_histo_b200.net as net
    from ssl_cr_histo_b200 import _lib, ddp, losses, optim

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus and world > 1:
        raise SystemExit("WORLD_SIZE %d != --gpus %d" % (world, args.gpus))
    if args.gpus > 1 and world == 1:
        raise SystemExit("launch with torchrun --nproc-per-node %d for --gpus %d" % (args.gpus, args.gpus))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if _lib.load().b2n_device_ok() != 1:
        raise SystemExit("bench needs a compute-capability 10.x GPU (no fallback path exists)")
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    b, mu, S = args.batch_size, args.mu, args.size
    nx, nu = 3 * b, b * mu
    torch.manual_seed(42)
    if args.workload == "rsp":
        return run_rsp(args, net, losses, optim, ddp, dist, dev, rank, world, local)
    student, cls_s = net.TripletNet_Finetune("resnet18"), net.FinetuneResNet(1)
    import copy
    teacher, cls_t = copy.deepcopy(student), copy.deepcopy(cls_s)
    for p in list(teacher.parameters()) + list(cls_t.parameters()):   # --modules_teacher 64 (:414-427)
        p.requires_grad = False
    student, cls_s, teacher, cls_t = (m.to(dev) for m in (student, cls_s, teacher, cls_t))
    teacher.eval(); cls_t.eval(); student.train(); cls_s.train()
    params = list(student.parameters()) + list(cls_s.parameters())
    # eval_BreastPathQ_SSL_CR.py:481 -- Adam(lr 1e-4, wd 1e-4); the multi-tensor drop-in by default
    if args.torch_optim:
        opt = torch.optim.Adam(params, lr=1e-4, betas=(0.9, 0.999), weight_decay=1e-4)
    else:
        opt = optim.Adam(params, lr=1e-4, betas=(0.9, 0.999), weight_decay=1e-4)
        opt.grad_scale = 1.0 / world          # folds the all-reduce averaging into the step
    reducer = ddp.GradAllReducer(params) if world > 1 else None

    # synthetic inputs: host copies in pinned memory (e2e) and resident device copies (value)
    seed = 1000 * rank
    def patches_u8(n, sd):   # uint8-valued floats, un-normalised (dataset.py:65-67)
        g = torch.Generator().manual_seed(sd)
        return torch.randint(0, 256, (n, 3, S, S), dtype=torch.uint8, generator=g).float()

    host = [patches_u8(n, seed + i) for i, n in enumerate((nx, nu, nu))]
    host_t = torch.rand(nx, generator=torch.Generator().manual_seed(seed + 7)).pin_memory()
    # `value`: fp32 patches resident in HBM, exactly what the reference's loop holds after its
    # .float() / .cuda() (eval_BreastPathQ_SSL_CR.py:68-71)
    resident = [h.to(dev) for h in host] + [host_t.to(dev)]
    # `e2e`: the patches travel as the uint8 pixels the dataset holds (dataset.py:65-67); the
    # trunk's stem pack kernel does the cast on the device (4x fewer PCIe bytes than shipping the
    # loop's fp32 copy).  --e2e-fp32 ships fp32 instead.
    if args.e2e_fp32:
        host = [h.pin_memory() for h in host]
    else:
        host = [h.to(torch.uint8).pin_memory() for h in host]
    h2d_bytes = sum(h.numel() * h.element_size() for h in host) + host_t.numel() * 4

    def step(ix, iw, is_, tx):
        with torch.no_grad():
            logits_u_w = cls_t(teacher(iw))
        logits = cls_s(student(torch.cat((ix, is_))))
        loss, parts = losses.consistency_mse(logits[:nx], tx, logits_u_w, logits[nx:], 1.0)
        if reducer is not None:
            reducer.zero_grad()
        else:
            opt.zero_grad(set_to_none=True)
        loss.backward()
        if reducer is not None:
            reducer.all_reduce(average=args.torch_optim)
        opt.step()
        return loss

    # e2e: every step's inputs travel pinned host -> device inside the timed region.  Like a
    # DataLoader with pin_memory (pretrain_BreastPathQ.py:213) feeding `.cuda(non_blocking=True)`,
    # the copy of step i+1 is issued on a side stream while step i computes (two device-side
    # input slots); the loss of every step is read back on the host (:103).
    copy_stream = torch.cuda.Stream(device=dev)
    slots = [[torch.empty(h.shape, dtype=h.dtype, device=dev) for h in host + [host_t]] for _ in range(2)]
    ready = [torch.cuda.Event(), torch.cuda.Event()]
    consumed = [torch.cuda.Event(), torch.cuda.Event()]
    e2e_state = {"i": 0, "primed": False}

    def issue_copy(slot):
        with torch.cuda.stream(copy_stream):
            copy_stream.wait_event(consumed[slot])
            for d, h in zip(slots[slot], host + [host_t]):
                d.copy_(h, non_blocking=True)
            ready[slot].record(copy_stream)

    def step_e2e():
        cur = e2e_state["i"] & 1
        if not e2e_state["primed"]:
            consumed[0].record(); consumed[1].record()
            issue_copy(cur)
            e2e_state["primed"] = True
        issue_copy(cur ^ 1)                               # prefetch the next step's batch
        torch.cuda.current_stream().wait_event(ready[cur])
        loss = step(*slots[cur])
        consumed[cur].record()
        e2e_state["i"] += 1
        return float(loss.detach())                       # the loop's loss.item() read-back

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        n0 = _lib.launch_count()
        e0.record()
        for _ in range(steps):
            fn()
        e1.record()
        barrier()
        ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms), _lib.launch_count() - n0

    for _ in range(max(args.warmup, 3)):
        step(*resident)
    sampler = ClockSampler(local)
    sampler.start()
    ms, launches = timed(lambda: step(*resident), args.steps)
    sampler.stop_flag = True
    step_e2e()
    ms_e2e, _ = timed(step_e2e, args.steps)

    # per-kernel roofline of the dominant kernels, CUDA events on the launching stream
    roof = None
    if not args.no_profile:
        _lib.PROFILE = {"b2n_conv_fwd": [], "b2n_conv_wgrad": []}
        t_ms, _ = timed(lambda: step(*resident), 2)
        prof, _lib.PROFILE = _lib.PROFILE, None
        torch.cuda.synchronize()
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        bf16 = peaks.get("bf16_tflops_sustained")
        peak_src = ("MEASURED_PEAKS.json bf16_tflops_sustained (kernels timed inside a long step); "
                    "TF32 = half of it (TF32 issues at half the bf16 rate, the file has no TF32 line)")
        if bf16 is None:
            bf16, peak_src = 1400.0, "fallback 1.4 PFLOP/s sustained bf16 (B200_PROFILING.md); TF32 = half"
        peak16, peak32 = float(bf16), float(bf16) / 2.0
        if os.environ.get("B2N_PROF_DUMP"):
            with open(os.environ["B2N_PROF_DUMP"], "w") as f:
                json.dump({n: [(round(a.elapsed_time(c) * 1e3, 1), w[0]) for a, c, w in ev]
                           for n, ev in prof.items()}, f)
        # per group (forward convs / data gradients / weight gradients): time, algorithmic TFLOP/s,
        # and tensor-pipe utilisation = executed MMA FLOPs of each kind / that kind's peak
        grp = {}
        for name, ev in prof.items():
            for a, c, w in ev:
                g = grp.setdefault(w[3], {"ms": 0.0, "alg": 0.0, "pipe_s": 0.0, "n": 0, "bytes": 0.0})
                g["bytes"] += w[4]
                g["ms"] += a.elapsed_time(c)
                g["alg"] += w[0]
                g["pipe_s"] += w[1] / (peak16 * 1e12) + w[2] / (peak32 * 1e12)
                g["n"] += 1
        step_ms = ms / args.steps   # the timed run (the profiled pass is slowed by its own events)

        def summary(keys):
            ms = sum(grp[k]["ms"] for k in keys)
            alg = sum(grp[k]["alg"] for k in keys)
            pipe = sum(grp[k]["pipe_s"] for k in keys)
            return {"launches_per_step": sum(grp[k]["n"] for k in keys) / 2, "ms_per_step": ms / 2,
                    "algorithmic_dram_bytes_per_step": sum(grp[k]["bytes"] for k in keys) / 2,
                    "algorithmic_tflops": alg / (ms * 1e-3) / 1e12,
                    "tensor_pipe_util": pipe / (ms * 1e-3), "share_of_step": (ms / 2) / step_ms}

        dom = summary(["fwd", "dgrad"])
        # measured DRAM traffic of the same launches: one ncu pass over a step of this exact
        # configuration, committed under profiles/ (bytes per step, summed over the launches)
        traffic, traffic_src = None, None
        try:
            tr = json.load(open(os.path.join(ROOT, "profiles", "r1_conv_traffic.json")))
            if tr["config"] == {"batch_size": b, "mu": mu, "size": S}:
                k = tr["per_step"]["conv_igemm_kernel"]
                traffic = k["dram_read_bytes"] + k["dram_write_bytes"]
                traffic_src = "profiles/r1_conv_traffic.json (ncu dram__bytes_read+write, per step, %d launches)" % k["launches"]
        except Exception:
            pass
        roof = {"bound": "tensor", "kernel": "conv_igemm_kernel (forward + data-gradient launches)",
                "achieved": dom["algorithmic_tflops"], "peak": peak32, "unit": "TFLOP/s",
                "frac": dom["algorithmic_tflops"] / peak32, "traffic": traffic,
                "traffic_source": traffic_src,
                "algorithmic_dram_bytes_per_step": dom["algorithmic_dram_bytes_per_step"],
                "peak_source": peak_src, "share_of_step": dom["share_of_step"],
                "launches_per_step": dom["launches_per_step"],
                # executed MMA math / peak of the MMA kind: the forward issues 3 FP16 MMAs per
                # product (error compensation), which `frac` (algorithmic FLOPs vs the TF32 peak)
                # counts once
                "tensor_pipe_util": dom["tensor_pipe_util"],
                "forward": summary(["fwd"]), "dgrad": summary(["dgrad"]), "wgrad": summary(["wgrad"]),
                "peaks_tflops": {"fp16_mma": peak16, "tf32_mma": peak32}}

    patches = (nx + nu) * world
    alg_flops = world * (nu * FLOP_FWD + (nx + nu) * (FLOP_FWD + FLOP_BWD))
    line = {
        "metric": "224x224 histo patches/sec (consistency step)", "unit": "patches/s",
        "value": patches * args.steps / (ms * 1e-3), "n_gpus": world, "steps": args.steps,
        "warmup": max(args.warmup, 3), "ms_per_step": ms / args.steps, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None,
        "dtype": "fp16 (hi,lo) error-compensated forward MMAs + tf32 backward MMAs, fp32 accumulate / storage",
        "data": "synthetic",
        "config": {"workload": "SSL_CR consistency step (eval_BreastPathQ_SSL_CR.py:76-100), MSE/MSE, "
                               "modules_student=0, BASELINE configs[2] per rank",
                   "labeled": nx, "unlabeled_weak": nu, "unlabeled_strong": nu, "image": S,
                   "unique_patches_per_step_per_rank": nx + nu, "optimizer": "Adam(1e-4, wd 1e-4), " + ("torch.optim" if args.torch_optim else "multi-tensor kernel"),
                   "parallelism": "dp%d, one NCCL all-reduce of %.1f MB grads/step" % (
                       world, reducer.nbytes / 1e6) if reducer else "single GPU",
                   "l2": "inputs (%.0f MB/step fp32) larger than the 126 MB L2"
                         % (sum(r.numel() * 4 for r in resident) / 1e6),
                   "e2e_input_format": "fp32 NCHW" if args.e2e_fp32 else "uint8 NCHW (cast in the stem kernel)"},
        "algorithmic_tflops": alg_flops * args.steps / (ms * 1e-3) / 1e12,
        "e2e": {"value": patches * args.steps / (ms_e2e * 1e-3), "unit": "patches/s",
                "h2d_bytes_per_step": h2d_bytes, "d2h_bytes_per_step": 4,
                "ms_per_step": ms_e2e / args.steps},
        "gpu_launches": launches,
        "clocks": sampler.summary(),
    }
    if roof is not None:
        line["roofline"] = roof
    if rank == 0 and not args.no_cpu_baseline and world == 1:
        r = cpu_reference_step_rate(3, 1, size=S)
        line["cpu_baseline"] = {"value": r["value"], "unit": "patches/s", "cores": r["cores"],
                                "kind": "port", "sample": r["sample"]}
    if rank == 0:
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def run_rsp(args, net, losses, optim, ddp, dist, dev, rank, world, local):
    """BASELINE configs[1]: the RSP pretext step of pretrain_BreastPathQ.py:42-68 -- TripletNet over a
    resolution triple (three trunk passes, shared weights, per-pass BN statistics), Classifier(768,6),
    cross-entropy over the 6 orders, SGD-Nesterov(lr .01, momentum .9, wd 1e-4).  One step = 3 * batch
    patches.  Secondary workload: prints the same JSON line without roofline / cpu_baseline."""
    from ssl_cr_histo_b200 import _lib
    nb, S = args.rsp_batch, args.size
    model, cls = net.TripletNet("resnet18").to(dev).train(), net.Classifier(768, 6).to(dev).train()
    params = list(model.parameters()) + list(cls.parameters())
    opt = (torch.optim.SGD if args.torch_optim else optim.SGD)(
        params, lr=0.01, momentum=0.9, weight_decay=1e-4, nesterov=True)       # :245
    if not args.torch_optim:
        opt.grad_scale = 1.0 / world
    reducer = ddp.GradAllReducer(params) if world > 1 else None
    g = torch.Generator().manual_seed(1000 * rank)
    host = [torch.randint(0, 256, (nb, 3, S, S), dtype=torch.uint8, generator=g).pin_memory() for _ in range(3)]
    host_t = torch.randint(0, 6, (nb,), generator=g).pin_memory()
    resident = [h.to(dev).float() for h in host] + [host_t.to(dev)]

    def step(i1, i2, i3, target):
        loss, pred = losses.cross_entropy(cls(model(i1, i2, i3)), target)       # :54-56,66
        if reducer is not None:
            reducer.zero_grad()
        else:
            opt.zero_grad(set_to_none=True)
        loss.backward()
        if reducer is not None:
            reducer.all_reduce(average=args.torch_optim)
        opt.step()
        return loss

    def step_e2e():
        dev_in = [h.to(dev, non_blocking=True) for h in host] + [host_t.to(dev, non_blocking=True)]
        return float(step(*dev_in).detach())

    def timed(fn, steps):
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        n0 = _lib.launch_count()
        e0.record()
        for _ in range(steps):
            fn()
        e1.record()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms), _lib.launch_count() - n0

    for _ in range(max(args.warmup, 3)):
        step(*resident)
    sampler = ClockSampler(local)
    sampler.start()
    ms, launches = timed(lambda: step(*resident), args.steps)
    sampler.stop_flag = True
    step_e2e()
    ms_e2e, _ = timed(step_e2e, args.steps)
    patches = 3 * nb * world
    line = {
        "metric": "224x224 histo patches/sec (RSP pretext step)", "unit": "patches/s",
        "value": patches * args.steps / (ms * 1e-3), "n_gpus": world, "steps": args.steps,
        "warmup": max(args.warmup, 3), "ms_per_step": ms / args.steps, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None,
        "dtype": "fp16 (hi,lo) error-compensated forward MMAs + tf32 backward MMAs, fp32 accumulate / storage",
        "data": "synthetic",
        "config": {"workload": "RSP pretext step (pretrain_BreastPathQ.py:42-68), BASELINE configs[1] per rank",
                   "triples": nb, "image": S, "patches_per_step_per_rank": 3 * nb,
                   "optimizer": "SGD-Nesterov(.01, .9, wd 1e-4), " + ("torch.optim" if args.torch_optim else "multi-tensor kernel"),
                   "l2": "inputs (%.0f MB/step fp32) larger than the 126 MB L2" % (3 * nb * 3 * S * S * 4 / 1e6)},
        "algorithmic_tflops": patches * (FLOP_FWD + FLOP_BWD) * args.steps / (ms * 1e-3) / 1e12,
        "e2e": {"value": patches * args.steps / (ms_e2e * 1e-3), "unit": "patches/s",
                "h2d_bytes_per_step": sum(h.numel() for h in host) + host_t.numel() * 8,
                "d2h_bytes_per_step": 4, "ms_per_step": ms_e2e / args.steps},
        "gpu_launches": launches, "clocks": sampler.summary(),
    }
    if rank == 0:
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    a = parse()
    if a.impl == "reference":
        run_reference(a)
    else:
        run_b200(a)
