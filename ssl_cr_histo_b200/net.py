"""Drop-in replacements for the reference's ``models/net.py`` classes.

Same constructor and ``forward`` signatures, same parameter / buffer names and enumeration
order (so ``load_state_dict`` of reference checkpoints, index-based freezing
(eval_BreastPathQ_SSL_CR.py:414-441), ``copy.deepcopy`` teacher hand-off and ``torch.optim``
all work unchanged) -- but every FLOP runs in the hand-written sm_100a kernels of libb2n.so.

    import ssl_cr_histo_b200.net as net           # instead of `import models.net as net`
    model = net.TripletNet('resnet18'); classifier = net.Classifier(768, 6)
"""
from __future__ import annotations

from torch import nn

from . import heads
from .trunk import ResNet18Trunk


def _pair_mlp() -> nn.Sequential:
    # parameter container only (models/net.py:36-37); evaluated by heads.mlp2
    return nn.Sequential(nn.Linear(512 * 2, 512), nn.ReLU(True), nn.Linear(512, 256))


class Classifier(nn.Module):
    """models/net.py:8-20 -- Linear(in,128) + ReLU + Linear(128,num_classes)."""

    def __init__(self, in_features, num_classes):
        super(Classifier, self).__init__()
        self.classifier = nn.Sequential(
            nn.Linear(in_features, 128),
            nn.ReLU(True),
            nn.Linear(128, num_classes))

    def forward(self, x):
        return heads.mlp2(x, self.classifier[0], self.classifier[2])


class TripletNet(nn.Module):
    """models/net.py:25-66 -- siamese ResNet18 over a resolution triple, pairwise MLP."""

    def __init__(self, model):
        super(TripletNet, self).__init__()
        if model == 'resnet18':
            self.model = ResNet18Trunk()
            self.fc = _pair_mlp()
        else:
            # the reference also has a resnet50 branch (models/net.py:39-45) that none of its
            # scripts uses; it is outside this package's scope
            raise NotImplementedError('not supported model type: {}'.format(model))

    def forward(self, i1, i2, i3):
        # three passes with shared weights; BN batch statistics per pass, running statistics
        # updated in the order i1, i2, i3 exactly as models/net.py:51-53
        E1 = self.model(i1)
        E2 = self.model(i2)
        E3 = self.model(i3)
        # pair concat + fc x3 + concat (:56-64) without materialising any concatenation
        return heads.pair_mlp(E1, E2, E3, self.fc[0], self.fc[2])


class TripletNet_Finetune(nn.Module):
    """models/net.py:70-103.  The reference runs the trunk three times on the *same* input;
    here it runs once: E1 == E2 == E3 bit-identically, autograd sums the three upstream
    gradients, and the BatchNorm running statistics receive the three identical updates in
    closed form (``n_updates=3``), so outputs, gradients and buffers match the reference."""

    def __init__(self, model):
        super(TripletNet_Finetune, self).__init__()
        if model == 'resnet18':
            self.model = ResNet18Trunk()
            self.fc = _pair_mlp()
        else:
            raise NotImplementedError('not supported model type: {}'.format(model))

    def forward(self, i):
        E = self.model(i, n_updates=3)
        return heads.pair_mlp_same(E, self.fc[0], self.fc[2])


class FinetuneResNet(nn.Module):
    """models/net.py:107-115 -- Linear(768, num_classes)."""

    def __init__(self, num_classes):
        super(FinetuneResNet, self).__init__()
        self.classifier = nn.Sequential(nn.Linear(256 * 3, num_classes))

    def forward(self, x):
        return heads.linear(x, self.classifier[0])
