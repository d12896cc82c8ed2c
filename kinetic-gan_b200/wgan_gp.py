"""WGAN-GP training step with the reference's semantics (kinetic-gan.py:94-114 gradient penalty,
:137-174 loop body, Adam at :77-78), restructured for B200:

  * parameters, gradients and Adam moments of each network live in ONE flat fp32 buffer, so the optimizer is a
    single fused launch (kgan_adam_step) and the DDP gradient exchange is a single all-reduce (ddp.py);
  * the critic step runs G under no_grad and the generator step skips the critic's weight gradients: both are
    zeroed before any optimizer reads them in the reference (kinetic-gan.py:137,157), so results are identical
    (SURVEY.md §7 I6);
  * no per-iteration host synchronisation: losses stay on the device unless asked for.
"""
import numpy as np
import torch

from . import ops


def compute_gradient_penalty(D, real_samples, fake_samples, labels, alpha=None, return_gradients=False):
    """Calculates the gradient penalty loss for WGAN GP (kinetic-gan.py:94-114).  `alpha` (N,1,1,1) defaults to the
    reference's host draw np.random.random at :97."""
    n = real_samples.size(0)
    if alpha is None:
        alpha = torch.as_tensor(np.random.random((n, 1, 1, 1)), dtype=real_samples.dtype, device=real_samples.device)
    interpolates = ops.interpolate(alpha.reshape(n).contiguous(), real_samples.contiguous(),
                                   fake_samples.contiguous()).requires_grad_(True)
    d_interpolates = D(interpolates, labels)
    fake = torch.ones(n, 1, dtype=real_samples.dtype, device=real_samples.device)
    gradients = torch.autograd.grad(outputs=d_interpolates, inputs=interpolates, grad_outputs=fake, create_graph=True,
                                    retain_graph=True, only_inputs=True)[0]
    flat = gradients.reshape(n, -1)
    gradient_penalty = ((flat.norm(2, dim=1) - 1) ** 2).mean()
    return (gradient_penalty, gradients) if return_gradients else gradient_penalty
