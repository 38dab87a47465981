"""FITC -- host mirror of the reference's sparse strategy `FITC(x, Xu, y, mean, kern, logNoise)`
(src/sparse/fully_indep_train_conditional.jl:335-338 == GPE(..., FullyIndepStrat(Xu))), over the
gpb200_fitc_* entry points: update_cK!, alpha / mll / logdet, dmll_noise, dmll_mean!, dmll_kern!,
predict_f / predict_y."""
import math

import numpy as np

from . import capi
from .gpe import _as_dxn
from .kernels import flatten
from .means import MeanZero


class FITC:
    _mode = 0

    def __init__(self, x, Xu, y, mean=None, kernel=None, logNoise=-2.0, device=0, distributed=False):
        """distributed=True (one process per GPU, torch.distributed initialised): x / y are THIS rank's slice of the observations,
        Xu / kernel / logNoise are the same everywhere; mll and the gradients are those of the full data set, alpha is the
        slice's."""
        self.mean = mean if mean is not None else MeanZero()
        self.kernel = kernel
        self.logNoise = float(logNoise)
        self.x = _as_dxn(x)
        self.Xu = _as_dxn(Xu)                       # inducing points, dim x M (one point per column)
        self.y = np.asarray(y, dtype=np.float64).ravel()
        self.dim, self.nobs = self.x.shape
        if self.Xu.shape[0] != self.dim or self.y.size != self.nobs:
            raise ValueError("Input and output observations must have consistent dimensions.")
        self._xpm = np.ascontiguousarray(self.x.T)
        self._eng = capi.FitcEngine(device)
        if distributed:
            import torch.distributed as dist
            from .dist import broadcast_bytes
            uid = broadcast_bytes(capi.Engine.nccl_unique_id(None) if dist.get_rank() == 0 else None, src=0)
            self._eng.comm_init(dist.get_world_size(), dist.get_rank(), uid)
        self._eng.set_mode(self._mode)
        self._eng.set_data(self._xpm, np.ascontiguousarray(self.Xu.T))
        ops, dims, theta, exposed = flatten(kernel, self.dim)
        self._exposed = exposed
        self._eng.set_kernel(ops, dims, theta.size)
        self.alpha = None
        self.mll = float("nan")
        self.dmll = None
        self.update_mll()

    def update_mll(self):
        """update_mll! (src/GPE.jl:202-212) with cK::FullyIndepPDMat."""
        self._eng.factorize(flatten(self.kernel, self.dim)[2], self.logNoise)
        mu = self.mean.mean(self._xpm)
        self.alpha, self.mll, self.logdet = self._eng.mll(self.y - mu)
        self.target = self.mll
        return self

    def update_dmll_noise_mean(self):
        """[dmll_noise (fitc.jl:243-257); dmll_mean! (GPE.jl:282-288)] -- kernel part not built yet."""
        out = [self._eng.grad_noise()]
        if self.mean.num_params() > 0:
            out.extend(self.mean.grad_stack(self._xpm).T @ self.alpha)
        self.dmll = np.array(out)
        return self

    def update_dmll(self, noise=True, domean=True, kern=True):
        """update_dmll! (src/GPE.jl:298-324) with the FITC strategy: [noise; mean; kernel]."""
        out = []
        if noise:
            out.append(self._eng.grad_noise())                                   # fitc.jl:243-257
        if domean and self.mean.num_params() > 0:
            out.extend(self.mean.grad_stack(self._xpm).T @ self.alpha)           # GPE.jl:282-288
        if kern:
            out.extend(self._eng.grad_kernel()[self._exposed])                   # fitc.jl:200-234 + sor.jl:219-253
        self.dmll = np.array(out)
        return self

    def update_mll_and_dmll(self, **kw):
        self.update_mll()
        return self.update_dmll(**kw)

    def noise_variance(self):
        return math.exp(2.0 * self.logNoise)

    def predict_f(self, x, full_cov=False):
        x = _as_dxn(x)
        if x.shape[0] != self.dim:
            raise ValueError("Gaussian Process object and input observations do not have consistent dimensions")
        xs = np.ascontiguousarray(x.T)
        if full_cov:                                                   # predict_full -> predictMVN (fitc.jl:324-332)
            mu, cov = self._eng.predict_cov(xs)
            return mu + self.mean.mean(xs), cov
        mu, var = self._eng.predict(xs)
        return mu + self.mean.mean(xs), np.maximum(var, 0.0)          # GP.jl:75

    def predict_y(self, x, full_cov=False):
        mu, s2 = self.predict_f(x, full_cov=full_cov)
        if full_cov:
            return mu, s2 + self.noise_variance() * np.eye(s2.shape[0])       # GPE.jl:412 (ScalMat)
        return mu, s2 + self.noise_variance()


class DTC(FITC):
    """Deterministic Training Conditional (src/sparse/determ_train_conditional.jl): SoR likelihood, FITC-style
    predictive variance."""
    _mode = 1


class SoR(FITC):
    """Subset of Regressors (src/sparse/subsetofregressors.jl)."""
    _mode = 2
