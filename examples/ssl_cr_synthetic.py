#!/usr/bin/env python
"""The reference's consistency-training loop (eval_BreastPathQ_SSL_CR.py:37-128, 505-516) end to end
on synthetic patches, with every stage of the hot path on the GPU kernels of this package:

    raw uint8 patches --(augment.TransformFix: weak / strong views, dataset.py:663-677)-->
    teacher (eval, no_grad) on the weak view, student (train) on cat(labeled, strong)
    --(losses.consistency_mse, :92-95)--> backward --(optim.Adam, :481)--> step,
    teacher <- student hand-off at the end of every epoch (weights.teacher_handoff_, :515-516).

    python examples/ssl_cr_synthetic.py                      # one GPU
    torchrun --nproc-per-node 2 --master-addr 127.0.0.1 examples/ssl_cr_synthetic.py   # data parallel

Only the lines marked `# b2n` differ from the reference's script.
"""
import argparse
import copy
import os
import sys

import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import ssl_cr_histo_b200.net as net                                   # b2n: instead of `import models.net as net`
from ssl_cr_histo_b200 import augment, ddp, losses, optim, weights    # b2n


def main(epochs=2, steps_per_epoch=3, batch_size=4, mu=2, image_size=64, lambda_u=1.0, seed=42, log=print):
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    dev = torch.device("cuda", int(os.environ.get("LOCAL_RANK", "0")))
    torch.cuda.set_device(dev)
    if world > 1:
        torch.distributed.init_process_group("nccl", device_id=dev)
    torch.manual_seed(seed)
    model_student, classifier_student = net.TripletNet_Finetune("resnet18"), net.FinetuneResNet(1)   # :383-387
    model_teacher, classifier_teacher = copy.deepcopy(model_student), copy.deepcopy(classifier_student)
    for p in list(model_teacher.parameters()) + list(classifier_teacher.parameters()):               # :414-427
        p.requires_grad = False
    for m in (model_student, classifier_student, model_teacher, classifier_teacher):
        m.to(dev)
    params = list(model_student.parameters()) + list(classifier_student.parameters())
    optimizer = optim.Adam(params, lr=1e-4, betas=(0.9, 0.999), weight_decay=1e-4)                   # b2n (:481)
    reducer = ddp.GradAllReducer(params, overlap=True) if world > 1 else None                        # b2n (:474-477)
    if reducer is not None:
        optimizer.grad_scale = 1.0 / world
    views = augment.TransformFix(image_size, N=7, seed=seed + rank)                                  # b2n (dataset.py:663)
    g = torch.Generator().manual_seed(seed + 1000 * rank)
    history = []
    for epoch in range(epochs):
        model_teacher.eval(); classifier_teacher.eval(); model_student.train(); classifier_student.train()
        for _ in range(steps_per_epoch):
            # the loaders' output: labeled items x 3 views and raw unlabeled patches, uint8 (dataset.py:65-67, :520)
            inputs_x = torch.randint(0, 256, (3 * batch_size, 3, image_size, image_size), dtype=torch.uint8, generator=g)
            targets_x = torch.rand(3 * batch_size, generator=g)
            raw_u = torch.randint(0, 256, (batch_size * mu, 3, image_size + 8, image_size + 8), dtype=torch.uint8,
                                  generator=g)
            inputs_x, targets_x, raw_u = inputs_x.to(dev), targets_x.to(dev), raw_u.to(dev)
            inputs_u_w, inputs_u_s = views(raw_u)                                                    # b2n: on the GPU
            with torch.no_grad():                                                                    # :79-81
                logits_u_w = classifier_teacher(model_teacher(inputs_u_w))
            logits = classifier_student(model_student(torch.cat((inputs_x, inputs_u_s))))            # :84-87
            logits_x, logits_u_s = logits[:inputs_x.shape[0]], logits[inputs_x.shape[0]:]
            loss, (sup, cons, _) = losses.consistency_mse(logits_x, targets_x, logits_u_w, logits_u_s, lambda_u)  # b2n (:92-95)
            if reducer is not None:
                reducer.zero_grad()
            else:
                optimizer.zero_grad()
            loss.backward()                                                                          # :99
            if reducer is not None:
                reducer.all_reduce(average=False)
            optimizer.step()                                                                         # :100
            history.append((float(loss.detach()), float(sup), float(cons)))      # the loop's loss.item()
        weights.teacher_handoff_(model_teacher, model_student)                                       # b2n (:515-516)
        weights.teacher_handoff_(classifier_teacher, classifier_student)
        if rank == 0:
            log("epoch %d  loss %.4f  (supervised %.4f, consistency %.4f)" % ((epoch,) + history[-1]))
    if world > 1:
        torch.distributed.destroy_process_group()
    return history, model_student, model_teacher


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--epochs", type=int, default=2)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--batch-size", type=int, default=4)
    ap.add_argument("--mu", type=int, default=2)
    ap.add_argument("--size", type=int, default=64)
    a = ap.parse_args()
    main(a.epochs, a.steps, a.batch_size, a.mu, a.size)
