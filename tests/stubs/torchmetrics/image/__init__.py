import torch
import torch.nn.functional as F
from torch import nn


class PeakSignalNoiseRatio(nn.Module):
    def __init__(self, data_range=1.0, **kwargs):
        super().__init__()
        self.data_range = float(data_range)

    def forward(self, preds, target):
        mse = torch.mean((preds - target) ** 2)
        return 10.0 * torch.log10(self.data_range**2 / mse)


class StructuralSimilarityIndexMeasure(nn.Module):
    """torchmetrics SSIM with its defaults (gaussian_kernel=True, sigma=1.5, k1=0.01, k2=0.03): separable 11-tap
    Gaussian window, reflect padding of (kernel - 1) / 2 cropped again before the mean."""

    def __init__(self, data_range=1.0, kernel_size=11, sigma=1.5, k1=0.01, k2=0.03, **kwargs):
        super().__init__()
        self.data_range, self.kernel_size, self.sigma, self.k1, self.k2 = float(data_range), kernel_size, sigma, k1, k2

    def forward(self, preds, target):
        ks, pad = self.kernel_size, (self.kernel_size - 1) // 2
        C = preds.shape[1]
        d = torch.arange((1 - ks) / 2, (1 + ks) / 2, 1, dtype=preds.dtype, device=preds.device)
        g = torch.exp(-((d / self.sigma) ** 2) / 2)
        g = (g / g.sum())[None]
        kernel = (g.t() @ g).expand(C, 1, ks, ks)
        c1, c2 = (self.k1 * self.data_range) ** 2, (self.k2 * self.data_range) ** 2
        p = F.pad(preds, (pad, pad, pad, pad), mode="reflect")
        t = F.pad(target, (pad, pad, pad, pad), mode="reflect")
        stack = torch.cat((p, t, p * p, t * t, p * t))
        out = F.conv2d(stack, kernel, groups=C)
        mu_p, mu_t, pp, tt, pt = out.split(preds.shape[0])
        s_p, s_t, s_pt = pp - mu_p**2, tt - mu_t**2, pt - mu_p * mu_t
        ssim = ((2 * mu_p * mu_t + c1) * (2 * s_pt + c2)) / ((mu_p**2 + mu_t**2 + c1) * (s_p + s_t + c2))
        return ssim[..., pad:-pad, pad:-pad].reshape(ssim.shape[0], -1).mean(-1).mean()


class MultiScaleStructuralSimilarityIndexMeasure(StructuralSimilarityIndexMeasure):
    pass
