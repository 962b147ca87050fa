"""TEST INFRASTRUCTURE ONLY -- stand-in for tensorflow_addons.losses.SigmoidFocalCrossEntropy (see ../tensorflow/__init__.py
for what the shim pins).  Restated from the published tfa implementation (sigmoid_focal_crossentropy): ce =
K.binary_crossentropy(y_true, y_pred, from_logits); p = sigmoid(y_pred) if from_logits else y_pred; p_t = y p + (1-y)(1-p);
alpha_t = y alpha + (1-y)(1-alpha); loss = sum(alpha_t (1-p_t)^gamma ce, axis=-1); reduction AUTO -> mean of the rest."""
import types

import torch

EPS = 1e-7      # K.epsilon()


class SigmoidFocalCrossEntropy:
    def __init__(self, from_logits=False, alpha=0.25, gamma=2.0, reduction="auto", name=None):
        self.from_logits, self.alpha, self.gamma = from_logits, alpha, gamma

    def __call__(self, y_true, y_pred):
        y = torch.as_tensor(y_true).as_subclass(torch.Tensor).to(torch.float32)
        x = torch.as_tensor(y_pred).as_subclass(torch.Tensor).to(torch.float32)
        if self.from_logits:
            ce = torch.clamp(x, min=0) - x * y + torch.log1p(torch.exp(-torch.abs(x)))      # sigmoid_cross_entropy_with_logits
            p = torch.sigmoid(x)
        else:
            pc = torch.clamp(x, EPS, 1.0 - EPS)
            ce = -(y * torch.log(pc + EPS) + (1 - y) * torch.log(1 - pc + EPS))
            p = x
        p_t = y * p + (1 - y) * (1 - p)
        alpha_t = y * self.alpha + (1 - y) * (1 - self.alpha)
        loss = (alpha_t * torch.pow(1.0 - p_t, self.gamma) * ce).sum(dim=-1)
        return loss.mean()


losses = types.SimpleNamespace(SigmoidFocalCrossEntropy=SigmoidFocalCrossEntropy)
