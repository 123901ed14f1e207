"""Host-side logic of the mode steps that needs no GPU: the declared-weights (training-step) form of the loss pack."""
import pytest
import torch

from unsupervised_depth_opticalflow_egomotion_b200 import mode_steps


def test_step_grad_matrix_is_the_weighted_total_gradient():
    """rows fl(w_k / B): what d(sum_k w_k mean_b loss_k[b]) / d loss_k[b] is for an upstream gradient of 1 (train.py:211-215)"""
    keys = ("a", "b", "c")
    w = {"a": 0.15, "b": 10.0, "c": 0.01, "unused": 3.0}
    m = mode_steps.step_grad_matrix(keys, w, 8, torch.device("cpu"))
    assert m.shape == (3, 8) and m.is_contiguous()
    loss = torch.rand(3, 8, requires_grad=True)
    total = sum(w[k] * loss[i].mean() for i, k in enumerate(keys))
    g, = torch.autograd.grad(total, loss)
    assert torch.allclose(m, g, rtol=1e-7, atol=0)
    assert mode_steps.step_grad_matrix(keys, w, 8, torch.device("cpu")) is m          # cached: no host-to-device copy per step


def test_loss_pack_refuses_other_weights_than_declared():
    keys = ("loss_a", "loss_b")
    pack = mode_steps.LossPack(torch.zeros(2, 4), keys, {"placeholder": torch.zeros(2)}, step_weights={"loss_a": 1.0, "loss_b": 0.5, "x": 9.0})
    assert set(pack) == {"loss_a", "loss_b", "placeholder"} and pack.step_weights == {"loss_a": 1.0, "loss_b": 0.5}
    with pytest.raises(ValueError):
        pack.total({"loss_a": 1.0, "loss_b": 0.25})
