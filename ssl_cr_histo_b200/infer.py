"""Whole-slide heat-map inference: the loop body of the reference's
``test_Camelyon16.py:30-70`` on the eval-mode kernels -- BatchNorm folded into the conv epilogues,
no intermediate fp32 activations, one tiny launch for the softmax 'tumor' column.

    probs_map = infer.probability_map(model, classifier, test_loader, test_loader.dataset.mask.shape)
"""
from __future__ import annotations

from typing import Iterable, Sequence, Tuple

import numpy as np
import torch

from . import _lib
from ._lib import call


def tumor_probabilities(model: torch.nn.Module, classifier: torch.nn.Module,
                        patches: torch.Tensor) -> torch.Tensor:
    """softmax(classifier(model(patches)))[:, -1] (test_Camelyon16.py:52-58) as a device tensor.
    ``patches``: (N,3,H,W) fp32 or uint8 CUDA tensor.  Switches nothing: call ``.eval()`` first,
    as the reference's ``test()`` does (:33-34)."""
    _lib.require_device(patches, "patch batch")
    with torch.no_grad():
        logits = classifier(model(patches)).contiguous()
        out = torch.empty(logits.shape[0], device=logits.device, dtype=torch.float32)
        call("b2n_softmax_last", logits, out, logits.shape[0], logits.shape[1])
    return out


def probability_map(model: torch.nn.Module, classifier: torch.nn.Module,
                    batches: Iterable[Tuple[torch.Tensor, torch.Tensor, torch.Tensor]],
                    mask_shape: Sequence[int]) -> np.ndarray:
    """test_Camelyon16.py:30-70: every batch is (patches, x_mask, y_mask); the tumour probability of
    patch i lands at probs_map[x_mask[i], y_mask[i]]."""
    model.eval()
    classifier.eval()
    probs_map = np.zeros(tuple(mask_shape))
    for patches, x_mask, y_mask in batches:
        probs = tumor_probabilities(model, classifier, patches.cuda(non_blocking=True))
        probs_map[np.asarray(x_mask), np.asarray(y_mask)] = probs.cpu().numpy()
    return probs_map
