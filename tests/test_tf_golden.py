"""Parity against real TensorFlow, when a golden file produced by tools/export_tf_golden.py is present."""
import os

import numpy as np
import pytest
import torch

from oracle import tem_oracle as O

GOLD = os.path.join(os.path.dirname(__file__), "golden", "tf_golden_3d.npz")


@pytest.mark.skipif(not os.path.exists(GOLD), reason="no TF golden file (TensorFlow is not installable in the build image)")
def test_oracle_matches_tensorflow_golden():
    z = np.load(GOLD)
    g = [torch.tensor(z[f"w_g_{i}"]) for i in range(12)]
    with torch.no_grad():
        y = O.generator_forward(g, torch.tensor(z["real_x"]), 8, True).numpy()
    np.testing.assert_allclose(y, z["fake_y_inference"], rtol=1e-4, atol=1e-5)
    dy = [torch.tensor(z[f"w_dy_{i}"]) for i in range(10)]
    with torch.no_grad():
        lg = O.discriminator_forward(dy, torch.tensor(z["fake_y_inference"]), 8, True)
    np.testing.assert_allclose(lg.numpy(), z["logit_dy"], rtol=1e-4, atol=1e-5)
    assert abs(float(O.generator_loss(lg)) - float(z["gen_loss"])) < 1e-5
