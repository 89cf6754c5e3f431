"""ssl_cr_histo_b200 -- B200-native (sm_100a) hot path of srinidhiPY/SSL_CR_Histo.

Public surface (mirrors the reference's operator interface for this path):
  net.Classifier / TripletNet / TripletNet_Finetune / FinetuneResNet   (models/net.py)
  losses.cross_entropy / consistency_mse / consistency_ce               (fused loss kernels)
  weights.lerp_ / teacher_handoff_ / lookahead_pull_                    (multi-tensor lerp)
  ddp.GradAllReducer                                                    (one NCCL all-reduce / step)
  optim.Adam / optim.SGD                                                (multi-tensor optimizer steps)
  graph.GraphedStep                                                     (whole-step CUDA-graph replay)
  augment.*                                                             (weak / strong views on the GPU)
  infer.probability_map                                                 (WSI heat-map inference loop)
"""
from . import net  # noqa: F401
from . import losses, weights, ddp, optim, graph  # noqa: F401

__all__ = ["net", "losses", "weights", "ddp", "optim", "graph"]
