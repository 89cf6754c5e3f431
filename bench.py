#!/usr/bin/env python
"""Headline benchmark: 224x224 histo patches/sec of the SSL_CR teacher-student consistency step.

    python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path
    python bench.py --impl reference --steps K --warmup W     # the reference's CPU path (oracle port)
    torchrun --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...   (N > 1, one rank per GPU)

Headline workload (BASELINE.json configs[2], SURVEY.md section 8d "cfg3"; per rank, weak scaling -- 8
ranks are configs[3]): eval_BreastPathQ_SSL_CR.py:76-100 with --batch_size 64 --mu 8 --lambda_u 1
--modules_student 0: labeled inputs_x (192,3,224,224), weak / strong unlabeled (512,3,224,224)
each; frozen teacher (eval, no_grad) + FinetuneResNet(1) on the weak view, student (train) on
cat(labeled, strong), MSE/MSE consistency loss, full backward, Adam(lr 1e-4, wd 1e-4).
One step = 704 unique patches per rank.  Synthetic uint8-valued patches, seeded random-init
weights (the reference's constructors under torch.manual_seed(42)).

Prints ONE JSON line (rank 0).  `value`: inputs resident in HBM; `e2e`: the same step through the
public modules with per-step pinned-host -> device input copies and a loss read-back.  The line
also carries `roofline` (dominant kernel, peaks measured in this run), `cpu_baseline` (oracle port
on the host cores), `gpu_library_baseline` (the reference modules on this GPU through torch + cuDNN)
and `secondary` (BASELINE configs[1] RSP pretext step and configs[4] Kather fine-tune step).
"""
import argparse
import copy
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
DTYPE = "fp16 (hi,lo) error-compensated forward MMAs + tf32 backward MMAs, fp32 accumulate / storage"


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=8)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference", "library"],
                    help="b200: this repo's kernels; reference: the reference's CPU path (oracle port) on the "
                         "host cores; library: the reference's modules on the GPU through torch + cuDNN")
    ap.add_argument("--batch-size", type=int, default=64, help="labeled items per rank (x3 views)")
    ap.add_argument("--mu", type=int, default=8)
    ap.add_argument("--size", type=int, default=224)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-profile", action="store_true")
    ap.add_argument("--no-library-baseline", action="store_true")
    ap.add_argument("--no-secondary", action="store_true")
    ap.add_argument("--no-peaks", action="store_true", help="skip the cuBLAS TF32 / FP16 peak measurement")
    ap.add_argument("--workload", default="cr", choices=["cr", "rsp", "kather"],
                    help="cr: SSL_CR consistency step (BASELINE configs[2], the headline metric); rsp: "
                         "RSP pretext step (configs[1]); kather: Kather 9-class fine-tune step (configs[4])")
    ap.add_argument("--rsp-batch", type=int, default=256, help="triples per rank for the rsp workload")
    ap.add_argument("--kather-batch", type=int, default=256, help="patches per rank for the kather workload")
    ap.add_argument("--torch-optim", action="store_true",
                    help="step torch.optim instead of the multi-tensor ssl_cr_histo_b200.optim kernels")
    ap.add_argument("--e2e-fp32", action="store_true",
                    help="e2e arm ships fp32 patches (the reference loop's format) instead of uint8")
    ap.add_argument("--graph", default="on", choices=["on", "off"],
                    help="replay the step from a CUDA graph (ssl_cr_histo_b200.graph.GraphedStep)")
    ap.add_argument("--overlap", default="on", choices=["on", "off"],
                    help="N > 1: start each gradient bucket's all-reduce as soon as it is complete")
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
    """SM clock and throttle reasons of one GPU, sampled DURING the timed region: NVML through
    pynvml (sub-millisecond per sample) when available, else nvidia-smi (tens of ms per sample)."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")
    NAMES = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.rows, self.stop_flag = index, [], False
        self.nvml = None
        try:
            import pynvml
            pynvml.nvmlInit()
            # torch's device index counts CUDA_VISIBLE_DEVICES entries; NVML counts physical GPUs
            vis = os.environ.get("CUDA_VISIBLE_DEVICES")
            phys = int(vis.split(",")[index]) if vis and vis.split(",")[index].strip().isdigit() else index
            self.handle = pynvml.nvmlDeviceGetHandleByIndex(phys)
            self.max_sm = pynvml.nvmlDeviceGetMaxClockInfo(self.handle, pynvml.NVML_CLOCK_SM)
            self.nvml = pynvml
        except Exception:
            self.nvml = None

    def _sample_nvml(self):
        n = self.nvml
        sm = n.nvmlDeviceGetClockInfo(self.handle, n.NVML_CLOCK_SM)
        get = getattr(n, "nvmlDeviceGetCurrentClocksEventReasons", None) or \
            n.nvmlDeviceGetCurrentClocksThrottleReasons
        r = get(self.handle)
        bits = [getattr(n, "nvmlClocksEventReason" + k, getattr(n, "nvmlClocksThrottleReason" + k, d))
                for k, d in (("HwSlowdown", 0x8), ("HwThermalSlowdown", 0x40),
                             ("SwThermalSlowdown", 0x20), ("SwPowerCap", 0x4))]
        return [str(sm), str(self.max_sm), ""] + ["Active" if r & b else "Not Active" for b in bits]

    def run(self):
        while not self.stop_flag:
            try:
                if self.nvml is not None:
                    self.rows.append(self._sample_nvml())
                    time.sleep(0.002)
                    continue
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
        reasons = [n for i, n in enumerate(self.NAMES)
                   if any(len(r) > 3 + i and r[3 + i].lower().startswith("active") for r in self.rows)]
        mx = [int(r[1]) for r in self.rows if len(r) > 1 and r[1].isdigit()]
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": reasons, "samples": len(self.rows),
                "source": "nvml" if self.nvml is not None else "nvidia-smi"}


class Env:
    """Process-wide context of the CUDA arm."""

    def __init__(self, args):
        import torch.distributed as dist
        self.args, self.dist = args, dist
        self.rank = int(os.environ.get("RANK", "0"))
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        self.local = int(os.environ.get("LOCAL_RANK", "0"))
        if self.world != args.gpus and self.world > 1:
            raise SystemExit("WORLD_SIZE %d != --gpus %d" % (self.world, args.gpus))
        if args.gpus > 1 and self.world == 1:
            raise SystemExit("launch with torchrun --nproc-per-node %d for --gpus %d" % (args.gpus, args.gpus))
        torch.cuda.set_device(self.local)
        self.dev = torch.device("cuda", self.local)
        if self.world > 1:
            dist.init_process_group("nccl", device_id=self.dev)

    def barrier(self):
        if self.world > 1:
            self.dist.barrier()
        torch.cuda.synchronize()

    def timed(self, fn, steps):
        """K calls of fn between barrier + synchronize, CUDA events, max over ranks ->
        (milliseconds, kernels this library launched or replayed)."""
        from ssl_cr_histo_b200 import _lib
        self.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        n0 = _lib.launch_count()
        e0.record()
        for _ in range(steps):
            fn()
        e1.record()
        self.barrier()
        ms = torch.tensor([e0.elapsed_time(e1)], device=self.dev)
        if self.world > 1:
            self.dist.all_reduce(ms, op=self.dist.ReduceOp.MAX)
        return float(ms), _lib.launch_count() - n0


class Workload:
    """One training step of a reference loop: models, optimizer, synthetic host / device inputs and
    the step function (device tensors in, loss tensor out; forward + loss + backward + all-reduce +
    optimizer step)."""
    name = metric = ""

    def build(self, env):
        raise NotImplementedError

    # -- shared plumbing
    def finish_build(self, env, params, passes):
        from ssl_cr_histo_b200 import ddp
        args = env.args
        self.reducer = None
        if env.world > 1:
            self.reducer = ddp.GradAllReducer(params, overlap=args.overlap == "on", passes=passes)
        self.params = params

    def backward_and_step(self, env, loss):
        if self.reducer is not None:
            self.reducer.zero_grad()
        else:
            self.opt.zero_grad(set_to_none=True)
        loss.backward()
        if self.reducer is not None:
            self.reducer.all_reduce(average=env.args.torch_optim)
        self.opt.step()
        return loss


def _patches_u8(n, size, seed):
    g = torch.Generator().manual_seed(seed)
    return torch.randint(0, 256, (n, 3, size, size), dtype=torch.uint8, generator=g)


class ConsistencyStep(Workload):
    name = "cr"
    metric = "224x224 histo patches/sec (consistency step)"

    def build(self, env):
        import ssl_cr_histo_b200.net as net
        from ssl_cr_histo_b200 import losses, optim
        args, dev = env.args, env.dev
        b, mu, S = args.batch_size, args.mu, args.size
        self.nx, self.nu, self.S = 3 * b, b * mu, S
        torch.manual_seed(42)
        student, cls_s = net.TripletNet_Finetune("resnet18"), net.FinetuneResNet(1)
        teacher, cls_t = copy.deepcopy(student), copy.deepcopy(cls_s)
        for p in list(teacher.parameters()) + list(cls_t.parameters()):   # --modules_teacher 64 (:414-427)
            p.requires_grad = False
        self.student, self.cls_s, self.teacher, self.cls_t = (m.to(dev) for m in (student, cls_s, teacher, cls_t))
        self.teacher.eval(); self.cls_t.eval(); self.student.train(); self.cls_s.train()
        params = list(self.student.parameters()) + list(self.cls_s.parameters())
        # eval_BreastPathQ_SSL_CR.py:481 -- Adam(lr 1e-4, wd 1e-4); the multi-tensor drop-in by default
        if args.torch_optim:
            self.opt = torch.optim.Adam(params, lr=1e-4, betas=(0.9, 0.999), weight_decay=1e-4,
                                        capturable=args.graph == "on")
        else:
            self.opt = optim.Adam(params, lr=1e-4, betas=(0.9, 0.999), weight_decay=1e-4,
                                  capturable=args.graph == "on")
            self.opt.grad_scale = 1.0 / env.world      # folds the all-reduce averaging into the step
        self.finish_build(env, params, passes=1)
        seed = 1000 * env.rank
        self.host_u8 = [_patches_u8(n, S, seed + i) for i, n in enumerate((self.nx, self.nu, self.nu))]
        self.host_t = torch.rand(self.nx, generator=torch.Generator().manual_seed(seed + 7))
        self.losses = losses
        self.patches = self.nx + self.nu
        self.alg_flops = self.nu * FLOP_FWD + (self.nx + self.nu) * (FLOP_FWD + FLOP_BWD)
        self.workload = ("SSL_CR consistency step (eval_BreastPathQ_SSL_CR.py:76-100), MSE/MSE, "
                         "modules_student=0, BASELINE configs[2] per rank")
        self.extra_config = {"labeled": self.nx, "unlabeled_weak": self.nu, "unlabeled_strong": self.nu,
                             "optimizer": "Adam(1e-4, wd 1e-4), " + ("torch.optim" if args.torch_optim
                                                                     else "multi-tensor kernel")}

    def host_inputs(self, fp32):
        """Pinned host copies of one step's inputs, in the order step() takes them."""
        ims = [h.float() if fp32 else h for h in self.host_u8]
        return [t.pin_memory() for t in ims + [self.host_t]]

    def step(self, env, ix, iw, is_, tx):
        with torch.no_grad():
            logits_u_w = self.cls_t(self.teacher(iw))
        logits = self.cls_s(self.student(torch.cat((ix, is_))))
        loss, _ = self.losses.consistency_mse(logits[:self.nx], tx, logits_u_w, logits[self.nx:], 1.0)
        return self.backward_and_step(env, loss)


class RspStep(Workload):
    """BASELINE configs[1]: the RSP pretext step of pretrain_BreastPathQ.py:42-68 -- TripletNet over a
    resolution triple (three trunk passes, shared weights, per-pass BN statistics), Classifier(768,6),
    cross-entropy over the 6 orders, SGD-Nesterov(lr .01, momentum .9, wd 1e-4)."""
    name = "rsp"
    metric = "224x224 histo patches/sec (RSP pretext step)"

    def build(self, env):
        import ssl_cr_histo_b200.net as net
        from ssl_cr_histo_b200 import losses, optim
        args, dev = env.args, env.dev
        nb, S = args.rsp_batch, args.size
        self.nb, self.S = nb, S
        torch.manual_seed(42)
        self.model = net.TripletNet("resnet18").to(dev).train()
        self.cls = net.Classifier(768, 6).to(dev).train()
        params = list(self.model.parameters()) + list(self.cls.parameters())
        self.opt = (torch.optim.SGD if args.torch_optim else optim.SGD)(
            params, lr=0.01, momentum=0.9, weight_decay=1e-4, nesterov=True)       # :245
        if not args.torch_optim:
            self.opt.grad_scale = 1.0 / env.world
        self.finish_build(env, params, passes=3)
        seed = 1000 * env.rank
        self.host_u8 = [_patches_u8(nb, S, seed + i) for i in range(3)]
        self.host_t = torch.randint(0, 6, (nb,), generator=torch.Generator().manual_seed(seed + 7))
        self.losses = losses
        self.patches = 3 * nb
        self.alg_flops = 3 * nb * (FLOP_FWD + FLOP_BWD)
        self.workload = "RSP pretext step (pretrain_BreastPathQ.py:42-68), BASELINE configs[1] per rank"
        self.extra_config = {"triples": nb, "optimizer": "SGD-Nesterov(.01, .9, wd 1e-4), " + (
            "torch.optim" if args.torch_optim else "multi-tensor kernel")}

    def host_inputs(self, fp32):
        ims = [h.float() if fp32 else h for h in self.host_u8]
        return [t.pin_memory() for t in ims + [self.host_t]]

    def step(self, env, i1, i2, i3, target):
        loss, _pred = self.losses.cross_entropy(self.cls(self.model(i1, i2, i3)), target)   # :54-56,66
        return self.backward_and_step(env, loss)


class KatherStep(Workload):
    """BASELINE configs[4] per rank: the supervised fine-tune step of eval_Kather_SSL.py:51-79 --
    TripletNet_Finetune (train, --modules 0: nothing frozen) + FinetuneResNet(9), cross-entropy,
    Adam(lr 1e-5, wd 1e-4) (:410,419-421); 256 patches per rank, 2048 over 8 ranks."""
    name = "kather"
    metric = "224x224 histo patches/sec (Kather 9-class fine-tune step)"

    def build(self, env):
        import ssl_cr_histo_b200.net as net
        from ssl_cr_histo_b200 import losses, optim
        args, dev = env.args, env.dev
        n, S = args.kather_batch, args.size
        self.n, self.S = n, S
        torch.manual_seed(42)
        self.model = net.TripletNet_Finetune("resnet18").to(dev).train()
        self.cls = net.FinetuneResNet(9).to(dev).train()
        params = list(self.model.parameters()) + list(self.cls.parameters())
        if args.torch_optim:
            self.opt = torch.optim.Adam(params, lr=1e-5, weight_decay=1e-4, capturable=args.graph == "on")
        else:
            self.opt = optim.Adam(params, lr=1e-5, weight_decay=1e-4, capturable=args.graph == "on")
            self.opt.grad_scale = 1.0 / env.world
        self.finish_build(env, params, passes=1)
        seed = 1000 * env.rank
        self.host_u8 = [_patches_u8(n, S, seed)]
        self.host_t = torch.randint(0, 9, (n,), generator=torch.Generator().manual_seed(seed + 7))
        self.losses = losses
        self.patches = n
        self.alg_flops = n * (FLOP_FWD + FLOP_BWD)
        self.workload = "Kather 9-class fine-tune step (eval_Kather_SSL.py:51-79), BASELINE configs[4] per rank"
        self.extra_config = {"patches": n, "optimizer": "Adam(1e-5, wd 1e-4), " + (
            "torch.optim" if args.torch_optim else "multi-tensor kernel")}

    def host_inputs(self, fp32):
        ims = [h.float() if fp32 else h for h in self.host_u8]
        return [t.pin_memory() for t in ims + [self.host_t]]

    def step(self, env, x, target):
        loss, _pred = self.losses.cross_entropy(self.cls(self.model(x)), target)             # :63-67
        return self.backward_and_step(env, loss)


WORKLOADS = {"cr": ConsistencyStep, "rsp": RspStep, "kather": KatherStep}


def measure(env, wl, steps, warmup, sample_clocks=True):
    """Warm up, then time `value` (fp32 patches resident in HBM, what the reference's loop holds after
    its .float() / .cuda(), eval_BreastPathQ_SSL_CR.py:68-71) and `e2e` (per-step pinned-host ->
    device copies of uint8 patches on a prefetch stream + the loop's loss.item() read-back)."""
    from ssl_cr_histo_b200 import graph as b2n_graph
    args, dev = env.args, env.dev
    host_f32 = wl.host_inputs(fp32=True)
    resident = [h.to(dev) for h in host_f32]
    host = host_f32 if args.e2e_fp32 else wl.host_inputs(fp32=False)
    del host_f32
    h2d_bytes = sum(h.numel() * h.element_size() for h in host)
    use_graph, graph_note = args.graph == "on", None

    def eager(*inp):
        return wl.step(env, *inp)

    for _ in range(max(warmup, 3)):
        eager(*resident)
    runners = {}
    if use_graph:
        # One captured graph per input format (fp32 resident / uint8 shipped): forward, loss,
        # backward, gradient all-reduce and optimizer step replay without per-launch host work.
        try:
            runners["value"] = b2n_graph.GraphedStep(eager, resident, warmup=1)
            runners["e2e"] = runners["value"] if args.e2e_fp32 else \
                b2n_graph.GraphedStep(eager, [h.to(dev) for h in host], warmup=1)
            ok = 1
        except Exception as exc:          # stay measurable if capture is refused on this box
            ok, graph_note = 0, "capture failed (%s: %s); eager launches" % (type(exc).__name__, str(exc)[:120])
            torch.cuda.synchronize()
        flag = torch.tensor([ok], device=dev)
        if env.world > 1:
            env.dist.all_reduce(flag, op=env.dist.ReduceOp.MIN)
        use_graph = bool(int(flag))
    if use_graph:
        g_val, g_e2e = runners["value"], runners["e2e"]
        step_value = g_val.replay            # inputs already sit in the graph's static buffers
    else:
        step_value = lambda: eager(*resident)   # noqa: E731

    # e2e: every step's inputs travel pinned host -> device inside the timed region.  Like a
    # DataLoader with pin_memory (pretrain_BreastPathQ.py:213) feeding `.cuda(non_blocking=True)`,
    # the copy of step i+1 is issued on a side stream while step i computes (two device-side
    # input slots); the loss of every step is read back on the host (:103).
    copy_stream = torch.cuda.Stream(device=dev)
    slots = [[torch.empty(h.shape, dtype=h.dtype, device=dev) for h in host] for _ in range(2)]
    ready = [torch.cuda.Event(), torch.cuda.Event()]
    consumed = [torch.cuda.Event(), torch.cuda.Event()]
    state = {"i": 0, "primed": False}

    def issue_copy(slot):
        with torch.cuda.stream(copy_stream):
            copy_stream.wait_event(consumed[slot])
            for d, h in zip(slots[slot], host):
                d.copy_(h, non_blocking=True)
            ready[slot].record(copy_stream)

    def step_e2e():
        cur = state["i"] & 1
        if not state["primed"]:
            consumed[0].record(); consumed[1].record()
            issue_copy(cur)
            state["primed"] = True
        issue_copy(cur ^ 1)                               # prefetch the next step's batch
        torch.cuda.current_stream().wait_event(ready[cur])
        loss = g_e2e(*slots[cur]) if use_graph else eager(*slots[cur])
        consumed[cur].record()
        state["i"] += 1
        return float(loss.detach())                       # the loop's loss.item() read-back

    step_value()
    sampler = ClockSampler(env.local) if sample_clocks else None
    if sampler:
        sampler.start()
    ms, launches = env.timed(step_value, steps)
    if sampler:
        sampler.stop_flag = True
    step_e2e()
    ms_e2e, _ = env.timed(step_e2e, steps)
    if use_graph:
        # replayed launches are not seen by the library's counter: count the graph's kernel nodes of
        # this library from an eager pass instead
        from ssl_cr_histo_b200 import _lib
        n0 = _lib.launch_count()
        eager(*resident)
        launches = (_lib.launch_count() - n0) * steps
        torch.cuda.synchronize()
    patches = wl.patches * env.world
    out = {
        "metric": wl.metric, "unit": "patches/s", "value": patches * steps / (ms * 1e-3),
        "ms_per_step": ms / steps,
        "algorithmic_tflops": env.world * wl.alg_flops * steps / (ms * 1e-3) / 1e12,
        "e2e": {"value": patches * steps / (ms_e2e * 1e-3), "unit": "patches/s",
                "h2d_bytes_per_step": h2d_bytes, "d2h_bytes_per_step": 4, "ms_per_step": ms_e2e / steps},
        "gpu_launches": launches,
        "config": dict({"workload": wl.workload, "image": wl.S,
                        "unique_patches_per_step_per_rank": wl.patches,
                        "parallelism": ("dp%d, %.1f MB gradient arena all-reduced over NCCL per step (%s)"
                                        % (env.world, wl.reducer.nbytes / 1e6,
                                           "%d buckets overlapped with backward" % len(wl.reducer.buckets)
                                           if wl.reducer.overlap else "one collective after backward"))
                        if wl.reducer else "single GPU",
                        "l2": "inputs (%.0f MB/step fp32) larger than the 126 MB L2"
                              % (sum(r.numel() * 4 for r in resident) / 1e6),
                        "launch": ("whole step replayed from one CUDA graph (graph.GraphedStep)" if use_graph
                                   else "eager ctypes launches" + (" -- " + graph_note if graph_note else "")),
                        "e2e_input_format": "fp32 NCHW" if args.e2e_fp32 else "uint8 NCHW (cast in the stem kernel)"},
                       **wl.extra_config),
    }
    if sampler:
        out["clocks"] = sampler.summary()
    return out, resident, eager


def measure_peaks(dev):
    """Dense tensor-core peaks of this GPU measured the way MEASURED_PEAKS.json was (torch.matmul
    8192^3 through cuBLAS: best of 10 = burst, back to back for ~1.5 s = sustained), for the two MMA
    kinds the conv kernels issue: TF32 (data / weight gradients) and FP16 (forward)."""
    n = 8192
    out = {}
    old = torch.backends.cuda.matmul.allow_tf32
    try:
        for kind, dtype, tf32 in (("tf32", torch.float32, True), ("fp16", torch.float16, False)):
            torch.backends.cuda.matmul.allow_tf32 = tf32
            a = torch.randn(n, n, device=dev, dtype=dtype)
            b = torch.randn(n, n, device=dev, dtype=dtype)
            for _ in range(3):
                a @ b
            torch.cuda.synchronize()
            best = 1e9
            for _ in range(10):
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record(); a @ b; e1.record(); torch.cuda.synchronize()
                best = min(best, e0.elapsed_time(e1))
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            reps = max(10, int(1500.0 / best))
            e0.record()
            for _ in range(reps):
                a @ b
            e1.record(); torch.cuda.synchronize()
            out[kind] = {"burst_tflops": 2.0 * n ** 3 / (best * 1e-3) / 1e12,
                         "sustained_tflops": 2.0 * n ** 3 * reps / (e0.elapsed_time(e1) * 1e-3) / 1e12}
            del a, b
    finally:
        torch.backends.cuda.matmul.allow_tf32 = old
    out["how"] = "torch.matmul %d^3 (cuBLAS), 2*N^3 FLOP: best of 10 (burst), %s back to back (sustained)" % (n, "~1.5 s")
    return out


def profile_launches(env, eager, resident):
    """Per-launch CUDA events around every b2n_conv_fwd / b2n_conv_dgrad_s2 / b2n_conv_wgrad call of
    two eager steps."""
    from ssl_cr_histo_b200 import _lib
    _lib.PROFILE = {"b2n_conv_fwd": [], "b2n_conv_dgrad_s2": [], "b2n_conv_dgrad_s2_sc": [],
                    "b2n_conv_wgrad": []}
    env.timed(lambda: eager(*resident), 2)
    prof, _lib.PROFILE = _lib.PROFILE, None
    torch.cuda.synchronize()
    return {n: [(a.elapsed_time(c), w) for a, c, w in ev] for n, ev in prof.items()}


def roofline(env, prof, step_ms, peaks_live):
    args = env.args
    file_peaks = {}
    try:
        file_peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    if peaks_live:
        peak16, peak32 = peaks_live["fp16"]["sustained_tflops"], peaks_live["tf32"]["sustained_tflops"]
        peak_src = ("measured in this run: cuBLAS 8192^3 sustained, TF32 for the data / weight gradients, "
                    "FP16 for the forward (kernels are timed inside a long step)")
    else:
        bf16 = file_peaks.get("bf16_tflops_sustained")
        peak_src = "MEASURED_PEAKS.json bf16_tflops_sustained; TF32 taken as half of it (not measured: --no-peaks)"
        if bf16 is None:
            bf16, peak_src = 1400.0, "fallback 1.4 PFLOP/s sustained bf16 (B200_PROFILING.md); TF32 = half"
        peak16, peak32 = float(bf16), float(bf16) / 2.0
    if os.environ.get("B2N_PROF_DUMP"):
        with open(os.environ["B2N_PROF_DUMP"], "w") as f:
            json.dump({n: [(round(ms * 1e3, 1), w[0]) for ms, w in ev] for n, ev in prof.items()}, f)
    grp = {}
    for name, ev in prof.items():
        for ms, w in ev:
            g = grp.setdefault(w[3], {"ms": 0.0, "alg": 0.0, "f16": 0.0, "tf32": 0.0, "n": 0, "bytes": 0.0})
            g["bytes"] += w[4]
            g["ms"] += ms
            g["alg"] += w[0]
            g["f16"] += w[1]
            g["tf32"] += w[2]
            g["n"] += 1

    def summary(keys):
        ms = sum(grp[k]["ms"] for k in keys)
        alg = sum(grp[k]["alg"] for k in keys)
        # algorithmic FLOPs / peak of the MMA kind that executes them (the forward's FP16
        # launches against the FP16 peak, TF32 launches against the TF32 peak), and the executed-MMA
        # version of the same (the forward issues three FP16 MMAs per product)
        alg_s = sum(grp[k]["alg"] / ((peak16 if grp[k]["f16"] > 0 else peak32) * 1e12) for k in keys)
        pipe_s = sum(grp[k]["f16"] / (peak16 * 1e12) + grp[k]["tf32"] / (peak32 * 1e12) for k in keys)
        return {"launches_per_step": sum(grp[k]["n"] for k in keys) / 2, "ms_per_step": ms / 2,
                "algorithmic_dram_bytes_per_step": sum(grp[k]["bytes"] for k in keys) / 2,
                "algorithmic_tflops": alg / (ms * 1e-3) / 1e12,
                "frac_of_own_mma_kind_peak": alg_s / (ms * 1e-3),
                "tensor_pipe_util": pipe_s / (ms * 1e-3), "share_of_step": (ms / 2) / step_ms}

    dom = summary(["fwd", "dgrad"])
    traffic, traffic_src = None, None
    for fn in ("r2_conv_traffic.json", "r1_conv_traffic.json"):
        try:
            tr = json.load(open(os.path.join(ROOT, "profiles", fn)))
            if tr["config"] == {"batch_size": args.batch_size, "mu": args.mu, "size": args.size}:
                k = tr["per_step"]["conv_igemm_kernel"]
                traffic = k["dram_read_bytes"] + k["dram_write_bytes"]
                traffic_src = "profiles/%s (ncu dram__bytes_read+write, per step, %d launches)" % (fn, k["launches"])
                break
        except Exception:
            pass
    return {"bound": "tensor", "kernel": "conv_igemm_kernel (forward + data-gradient launches)",
            "achieved": dom["algorithmic_tflops"], "peak": peak32, "unit": "TFLOP/s",
            "frac": dom["algorithmic_tflops"] / peak32,
            "frac_note": "algorithmic conv FLOPs (counted once although the forward issues 3 FP16 MMAs per "
                         "product) / measured TF32 peak; frac_of_own_mma_kind_peak and tensor_pipe_util "
                         "give the per-MMA-kind and executed-MMA views",
            "frac_of_own_mma_kind_peak": dom["frac_of_own_mma_kind_peak"],
            "tensor_pipe_util": dom["tensor_pipe_util"],
            "traffic": traffic, "traffic_source": traffic_src,
            "algorithmic_dram_bytes_per_step": dom["algorithmic_dram_bytes_per_step"],
            "peak_source": peak_src, "share_of_step": dom["share_of_step"],
            "launches_per_step": dom["launches_per_step"],
            "forward": summary(["fwd"]), "dgrad": summary(["dgrad"]), "wgrad": summary(["wgrad"]),
            "peaks_tflops": {"fp16_mma": peak16, "tf32_mma": peak32},
            "peaks_measured": peaks_live,
            "peaks_file": {k: file_peaks.get(k) for k in ("bf16_tflops", "bf16_tflops_sustained", "hbm_gbs")}}


def gpu_library_baseline(args, dev, steps=3, warmup=2):
    """The reference's own modules (oracle restatement of models/net.py over torch.nn, i.e. ATen +
    cuDNN + cuBLAS) on this GPU with the reference's settings -- fp32 NCHW, cudnn.benchmark = True
    (eval_BreastPathQ_SSL_CR.py:478), TF32 allowed for cuDNN convolutions (torch's default) -- on
    the same cfg3 batch.  `three_pass` is the reference as written (TripletNet_Finetune runs the
    trunk three times, models/net.py:88-90); `single_pass` evaluates the trunk once per model call,
    the same algorithmic work this repo's kernels do.  Bench-only: never on the product path."""
    from oracle import ref_net as O
    b, mu, S = args.batch_size, args.mu, args.size
    nx, nu = 3 * b, b * mu
    old = (torch.backends.cudnn.benchmark, torch.backends.cudnn.allow_tf32)
    torch.backends.cudnn.benchmark, torch.backends.cudnn.allow_tf32 = True, True
    ix, iw, is_ = (_patches_u8(n, S, i).float().to(dev) for i, n in enumerate((nx, nu, nu)))
    tx = torch.rand(nx, device=dev)
    res = {}

    class SinglePass(O.TripletNet_Finetune):
        def forward(self, i):
            e = self.model(i)
            return O._pairwise_features(self.fc, e, e, e)

    try:
        for name, cls in (("three_pass", O.TripletNet_Finetune), ("single_pass", SinglePass)):
            torch.manual_seed(42)
            student, cls_s = cls("resnet18").to(dev), O.FinetuneResNet(1).to(dev)
            teacher, cls_t = copy.deepcopy(student), copy.deepcopy(cls_s)
            for p in list(teacher.parameters()) + list(cls_t.parameters()):
                p.requires_grad = False
            teacher.eval(); cls_t.eval(); student.train(); cls_s.train()
            opt = O.make_cr_optimizer(list(student.parameters()) + list(cls_s.parameters()))
            try:
                for _ in range(warmup):
                    O.consistency_step(teacher, student, cls_t, cls_s, opt, ix, tx, iw, is_, 1.0, "mse")
                torch.cuda.synchronize()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                for _ in range(steps):
                    O.consistency_step(teacher, student, cls_t, cls_s, opt, ix, tx, iw, is_, 1.0, "mse")
                e1.record()
                torch.cuda.synchronize()
                ms = e0.elapsed_time(e1) / steps
                res[name] = {"ms_per_step": ms, "value": (nx + nu) / (ms * 1e-3), "unit": "patches/s"}
            except RuntimeError as exc:
                res[name] = {"error": str(exc)[:160]}
            del student, teacher, cls_s, cls_t, opt
            torch.cuda.empty_cache()
    finally:
        torch.backends.cudnn.benchmark, torch.backends.cudnn.allow_tf32 = old
    res["how"] = ("oracle/ref_net.py modules (torch.nn -> ATen/cuDNN/cuBLAS) on this GPU, fp32 NCHW, "
                  "cudnn.benchmark, cuDNN TF32 allowed, torch.optim.Adam, inputs resident, %d timed steps "
                  "after %d warm-up; same cfg3 batch" % (steps, warmup))
    return res


def run_b200(args):
    from ssl_cr_histo_b200 import _lib

    env = Env(args)
    if _lib.load().b2n_device_ok() != 1:
        raise SystemExit("bench needs a compute-capability 10.x GPU (no fallback path exists)")
    wl = WORKLOADS[args.workload]()
    wl.build(env)
    res, resident, eager = measure(env, wl, args.steps, args.warmup)
    line = {
        "metric": res["metric"], "unit": "patches/s", "value": res["value"], "n_gpus": env.world,
        "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": res["ms_per_step"],
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": DTYPE,
        "data": "synthetic", "config": res["config"], "algorithmic_tflops": res["algorithmic_tflops"],
        "e2e": res["e2e"], "gpu_launches": res["gpu_launches"], "clocks": res.get("clocks"),
    }
    if not args.no_profile:
        # (the per-launch pass runs before the peak measurement: seconds of back-to-back cuBLAS GEMMs
        # leave the part power-capped, which would slow the profiled launches, not the peaks)
        prof = profile_launches(env, eager, resident)
        peaks_live = None if args.no_peaks else measure_peaks(env.dev)
        line["roofline"] = roofline(env, prof, res["ms_per_step"], peaks_live)
    del resident, eager, wl
    torch.cuda.empty_cache()
    if not args.no_secondary and args.workload == "cr":
        # the other two GPU configurations BASELINE.json names, same harness, shorter runs
        sec = {}
        for name in ("rsp", "kather"):
            w2 = WORKLOADS[name]()
            w2.build(env)
            r2, res2, eager2 = measure(env, w2, max(2, args.steps // 2), 3, sample_clocks=False)
            sec[name] = {k: r2[k] for k in ("metric", "unit", "value", "ms_per_step", "algorithmic_tflops",
                                            "e2e", "gpu_launches", "config")}
            del w2, res2, eager2
            torch.cuda.empty_cache()
        line["secondary"] = sec
    if env.rank == 0 and env.world == 1 and not args.no_library_baseline and args.workload == "cr":
        line["gpu_library_baseline"] = gpu_library_baseline(args, env.dev)
    if env.rank == 0 and not args.no_cpu_baseline and env.world == 1:
        r = cpu_reference_step_rate(3, 1, size=args.size)
        line["cpu_baseline"] = {"value": r["value"], "unit": "patches/s", "cores": r["cores"],
                                "kind": "port", "sample": r["sample"]}
    if env.rank == 0:
        print(json.dumps(line))
    if env.world > 1:
        env.dist.destroy_process_group()


def run_library(args):
    """The reference's own modules on cuda:0 through torch + cuDNN (see gpu_library_baseline) as a
    contract-shaped line of its own (single GPU; the reference's multi-GPU mode is nn.DataParallel)."""
    if int(os.environ.get("RANK", "0")) != 0:
        return
    dev = torch.device("cuda", 0)
    r = gpu_library_baseline(args, dev, steps=max(args.steps, 2), warmup=max(args.warmup, 2))
    v = r["three_pass"]
    print(json.dumps({
        "impl": "library", "metric": "224x224 histo patches/sec (consistency step)",
        "value": v.get("value"), "unit": "patches/s", "n_gpus": 1, "steps": max(args.steps, 2),
        "warmup": max(args.warmup, 2), "ms_per_step": v.get("ms_per_step"), "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "fp32 storage, cuDNN TF32 convolutions (torch default)",
        "data": "synthetic",
        "config": {"workload": "SSL_CR consistency step (eval_BreastPathQ_SSL_CR.py:76-100), reference modules as "
                               "written (three trunk passes per model call), BASELINE configs[2]", "how": r["how"]},
        "single_pass_variant": r["single_pass"], "gpu_launches": 0}))


if __name__ == "__main__":
    a = parse()
    if a.impl == "reference":
        run_reference(a)
    elif a.impl == "library":
        run_library(a)
    else:
        run_b200(a)
