"""`add_diffusion_noise(image_tensor, noise_step)` -- the VCD baseline's image corruption (reference
vcd_utils/vcd_add_noise.py:3-28), kept here because this repo's `vcd_utils` package shadows the reference's when the
drivers import it (llava_sampling.py:23).  Forward diffusion q(x_t | x_0) = sqrt(abar_t) x_0 + sqrt(1 - abar_t) eps with a
1000-step sigmoid beta schedule between 1e-5 and 5e-3; eps is ONE torch.randn_like draw from the global generator, so a
seeded caller gets the same tensor as from the reference."""
import torch

_STEPS = 1000
_BETA_LO, _BETA_HI = 1e-5, 0.5e-2


def _signal_noise_scales():
    betas = torch.sigmoid(torch.linspace(-6, 6, _STEPS)) * (_BETA_HI - _BETA_LO) + _BETA_LO
    abar = torch.cumprod(1 - betas, dim=0)
    return torch.sqrt(abar), torch.sqrt(1 - abar)


def add_diffusion_noise(image_tensor: torch.Tensor, noise_step) -> torch.Tensor:
    signal, sigma = _signal_noise_scales()
    x0 = image_tensor.clone()
    eps = torch.randn_like(x0)
    return signal[noise_step] * x0 + sigma[noise_step] * eps
