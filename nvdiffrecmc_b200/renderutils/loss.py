"""PyTorch twin of the fused image loss (use_python=True), same role as the reference's render/renderutils/loss.py."""
import torch


def tonemap_log_srgb(x):
    x = torch.log(torch.clamp(x, min=0, max=65535) + 1)
    return torch.where(x > 0.0031308, torch.clamp(x, min=0.0031308) ** (1.0 / 2.4) * 1.055 - 0.055, 12.92 * x)


def image_loss_fn(img, target, loss, tonemapper):
    if tonemapper == 'log_srgb':
        img, target = tonemap_log_srgb(img), tonemap_log_srgb(target)
    d = img - target
    eps = 0.01
    if loss == 'mse':
        return (d * d).mean()
    if loss == 'smape':
        return (d.abs() / (img.abs() + target.abs() + eps)).mean()
    if loss == 'relmse':
        return (d * d / (img * img + target * target + eps)).mean()
    if loss == 'n2n':
        return (d * d / (img.detach() ** 2 + eps)).mean()
    return d.abs().mean()
