"""EnvironmentLight with the reference's API (render/light.py:21-59, 98-101) minus the
nvdiffrast-based image IO: lat-long probe `base` [H,W,3], `update_pdf()` builds the pdf and the
row / column CDFs that optix_env_shade consumes verbatim (`_pdf`, `rows[:,0]`, `cols`).

A CUDA probe goes through libmcshade's `mcs_update_pdf` (two launches, csrc/light.cu); a CPU probe evaluates the
reference's own torch formulas (the `use_python` twin -- there is no silent fallback: a CUDA probe without the library raises)."""
import ctypes as C
import numpy as np
import torch
from . import _lib as L


def pixel_grid(width, height, device, center_x=0.5, center_y=0.5):
    # render/util.py:62-66
    y, x = torch.meshgrid((torch.arange(0, height, dtype=torch.float32, device=device) + center_y) / height,
                          (torch.arange(0, width, dtype=torch.float32, device=device) + center_x) / width, indexing='ij')
    return torch.stack((x, y), dim=-1)


class EnvironmentLight:
    LIGHT_MIN_RES = 16
    MIN_ROUGHNESS = 0.08
    MAX_ROUGHNESS = 0.5

    def __init__(self, base):
        self.mtx = None
        self.base = base
        self.pdf_scale = (self.base.shape[0] * self.base.shape[1]) / (2 * np.pi * np.pi)
        self.update_pdf()

    def xfm(self, mtx):
        self.mtx = mtx

    def parameters(self):
        return [self.base]

    def clone(self):
        return EnvironmentLight(self.base.clone().detach())

    def clamp_(self, min=None, max=None):
        self.base.clamp_(min, max)

    def update_pdf(self, use_python=False):
        # light.py:46-59
        if self.base.is_cuda and not use_python:
            H, W = self.base.shape[0], self.base.shape[1]
            base = self.base.detach()
            if base.dtype != torch.float32:
                raise TypeError("EnvironmentLight.base must be float32")
            dev = base.device
            self._pdf = torch.empty(H, W, dtype=torch.float32, device=dev)
            self.cols = torch.empty(H, W, dtype=torch.float32, device=dev)
            rows = torch.empty(H, dtype=torch.float32, device=dev)
            self._row_totals = torch.empty(H, dtype=torch.float64, device=dev)
            b = L.view_hwc(base)
            L.check(L.lib().mcs_update_pdf(C.byref(b), self._pdf.data_ptr(), rows.data_ptr(), self.cols.data_ptr(), self._row_totals.data_ptr(),
                                           L.stream_ptr()), "update_pdf")
            self.rows = rows[:, None].expand(H, W)       # the reference stores rows replicated [H,W]; call sites take rows[:,0]
            return
        with torch.no_grad():
            Y = pixel_grid(self.base.shape[1], self.base.shape[0], self.base.device)[..., 1]
            self._pdf = torch.max(self.base, dim=-1)[0] * torch.sin(Y * np.pi)   # sin(theta) for lat-long
            self._pdf = self._pdf / torch.sum(self._pdf)
            self.cols = torch.cumsum(self._pdf, dim=1)
            self.rows = torch.cumsum(self.cols[:, -1:].repeat([1, self.cols.shape[1]]), dim=0)
            self.cols = self.cols / torch.where(self.cols[:, -1:] > 0, self.cols[:, -1:], torch.ones_like(self.cols))
            self.rows = self.rows / torch.where(self.rows[-1:, :] > 0, self.rows[-1:, :], torch.ones_like(self.rows))


def create_trainable_env_rnd(base_res, scale=0.5, bias=0.25, device="cuda"):
    # light.py:98-101
    base = torch.rand(base_res, base_res, 3, dtype=torch.float32, device=device) * scale + bias
    return EnvironmentLight(base.clone().detach().requires_grad_(True))
