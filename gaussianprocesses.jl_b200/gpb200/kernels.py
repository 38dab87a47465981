"""Host-side mirror of the reference's kernel types (src/kernels/*.jl): same constructor names,
argument meaning (log-scale hyper-parameters) and get_params / set_params! order.  The objects hold
parameters only; all arithmetic happens on the device.  `flatten()` turns a kernel tree into the
post-order program of include/gpb200.h; `spec()` gives the neutral tuple form the test oracle reads.

Differences from Julia worth knowing: `Masked` / `FixedKernel` take 0-based indices here.
"""
import numpy as np

from .capi import OP, OP_STRIDE, MAX_OPS, MAX_THETA, MAX_DIMS


class Kernel:
    _names = ()

    def get_params(self):
        raise NotImplementedError

    def set_params(self, hyp):
        raise NotImplementedError

    def num_params(self):
        return len(self.get_params())

    def get_param_names(self):
        return list(self._names)

    def spec(self):
        raise NotImplementedError

    def __add__(self, other):          # Base.:+  sum_kernel.jl:71
        return SumKernel(self, other)

    def __mul__(self, other):          # Base.:*  prod_kernel.jl:71
        return ProdKernel(self, other)

    # program emission: append ops; return list of positions in the FULL theta vector this
    # kernel exposes as free parameters (get_params order)
    def _emit(self, ctx, dims):
        raise NotImplementedError


class _Leaf(Kernel):
    _op = None
    _tag = None

    def __init__(self, *hyp):
        self.hyp = [float(v) for v in hyp]

    def get_params(self):
        return list(self.hyp)

    def set_params(self, hyp):
        hyp = [float(v) for v in hyp]
        if len(hyp) != len(self.hyp):
            raise ValueError("%s has %d parameters, received %d" % (self._tag, len(self.hyp), len(hyp)))
        self.hyp = hyp

    def _extra(self):
        return 0

    def _ard_dims(self):
        return None

    def spec(self):
        return (self._tag, list(self.hyp))

    def _emit(self, ctx, dims):
        nd_req = self._ard_dims()
        if nd_req is not None and nd_req != len(dims):
            raise ValueError("%s: %d length scales for %d active dimensions" % (self._tag, nd_req, len(dims)))
        toff = len(ctx["theta"])
        doff = len(ctx["dims"])
        ctx["theta"].extend(self.hyp)
        ctx["dims"].extend(int(k) for k in dims)
        ctx["ops"].append([self._op, toff, len(self.hyp), doff, len(dims), self._extra()])
        return list(range(toff, toff + len(self.hyp)))


def _leaf(tag, op, names):
    def deco(cls):
        cls._tag, cls._op, cls._names = tag, OP[op], names
        return cls
    return deco


@_leaf("SEIso", "SE_ISO", ("ll", "lσ"))
class SEIso(_Leaf):                      # se_iso.jl:10  SEIso(ll, lσ)
    def __init__(self, ll, lsig):
        super().__init__(ll, lsig)


@_leaf("Mat12Iso", "MAT12_ISO", ("ll", "lσ"))
class Mat12Iso(_Leaf):                   # mat12_iso.jl:12
    def __init__(self, ll, lsig):
        super().__init__(ll, lsig)


@_leaf("Mat32Iso", "MAT32_ISO", ("ll", "lσ"))
class Mat32Iso(_Leaf):                   # mat32_iso.jl:12
    def __init__(self, ll, lsig):
        super().__init__(ll, lsig)


@_leaf("Mat52Iso", "MAT52_ISO", ("ll", "lσ"))
class Mat52Iso(_Leaf):                   # mat52_iso.jl:12
    def __init__(self, ll, lsig):
        super().__init__(ll, lsig)


@_leaf("RQIso", "RQ_ISO", ("ll", "lσ", "lα"))
class RQIso(_Leaf):                      # rq_iso.jl:12
    def __init__(self, ll, lsig, lalpha):
        super().__init__(ll, lsig, lalpha)


@_leaf("Periodic", "PERIODIC", ("ll", "lσ", "lp"))
class Periodic(_Leaf):                   # periodic.jl:12
    def __init__(self, ll, lsig, lp):
        super().__init__(ll, lsig, lp)


@_leaf("LinIso", "LIN_ISO", ("ll",))
class LinIso(_Leaf):                     # lin_iso.jl:12
    def __init__(self, ll):
        super().__init__(ll)


@_leaf("Noise", "NOISE", ("lσ",))
class Noise(_Leaf):                      # noise.jl:12
    def __init__(self, lsig):
        super().__init__(lsig)


@_leaf("Const", "CONST", ("lσ",))
class Const(_Leaf):                      # const.jl:10
    def __init__(self, lsig):
        super().__init__(lsig)


@_leaf("Poly", "POLY", ("lc", "lσ"))
class Poly(_Leaf):                       # poly.jl:12  Poly(lc, lσ, deg)
    def __init__(self, lc, lsig, deg):
        super().__init__(lc, lsig)
        self.deg = int(deg)

    def _extra(self):
        return self.deg

    def spec(self):
        return ("Poly", list(self.hyp), self.deg)


class _ArdLeaf(_Leaf):
    _ntail = 1

    def __init__(self, ll, *tail):
        ll = [float(v) for v in np.atleast_1d(ll)]
        super().__init__(*(ll + [float(t) for t in tail]))
        self._nll = len(ll)

    def _ard_dims(self):
        return self._nll

    def get_param_names(self):
        return ["ll_%d" % (i + 1) for i in range(self._nll)] + list(self._names)


@_leaf("SEArd", "SE_ARD", ("lσ",))
class SEArd(_ArdLeaf):                   # se_ard.jl:13  SEArd(ll::Vector, lσ)
    def __init__(self, ll, lsig):
        super().__init__(ll, lsig)


@_leaf("Mat12Ard", "MAT12_ARD", ("lσ",))
class Mat12Ard(_ArdLeaf):                # mat12_ard.jl:13
    def __init__(self, ll, lsig):
        super().__init__(ll, lsig)


@_leaf("Mat32Ard", "MAT32_ARD", ("lσ",))
class Mat32Ard(_ArdLeaf):                # mat32_ard.jl:13
    def __init__(self, ll, lsig):
        super().__init__(ll, lsig)


@_leaf("Mat52Ard", "MAT52_ARD", ("lσ",))
class Mat52Ard(_ArdLeaf):                # mat52_ard.jl:13
    def __init__(self, ll, lsig):
        super().__init__(ll, lsig)


@_leaf("RQArd", "RQ_ARD", ("lσ", "lα"))
class RQArd(_ArdLeaf):                   # rq_ard.jl:13  RQArd(ll::Vector, lσ, lα)
    def __init__(self, ll, lsig, lalpha):
        super().__init__(ll, lsig, lalpha)


@_leaf("LinArd", "LIN_ARD", ())
class LinArd(_ArdLeaf):                  # lin_ard.jl:12  LinArd(ll::Vector)
    def __init__(self, ll):
        super().__init__(ll)


def SE(ll, lsig):                        # se.jl:14-15
    return SEArd(ll, lsig) if np.ndim(ll) else SEIso(ll, lsig)


def RQ(ll, lsig, lalpha):                # rq.jl:14-15
    return RQArd(ll, lsig, lalpha) if np.ndim(ll) else RQIso(ll, lsig, lalpha)


def Matern(nu, ll, lsig):                # mat.jl:42-74
    iso = {0.5: Mat12Iso, 1.5: Mat32Iso, 2.5: Mat52Iso}
    ard = {0.5: Mat12Ard, 1.5: Mat32Ard, 2.5: Mat52Ard}
    if nu not in iso:
        raise ValueError("Only Matern 1/2, 3/2 and 5/2 are implementable")
    return (ard if np.ndim(ll) else iso)[nu](ll, lsig)


class _Pair(Kernel):                     # pair_kernel.jl:1-36
    _op = None
    _tag = None

    def __init__(self, kleft, kright):
        self.kleft, self.kright = kleft, kright

    def get_params(self):
        return self.kleft.get_params() + self.kright.get_params()

    def set_params(self, hyp):
        npl = self.kleft.num_params()
        self.kleft.set_params(hyp[:npl])
        self.kright.set_params(hyp[npl:])

    def get_param_names(self):
        return self.kleft.get_param_names() + self.kright.get_param_names()

    def spec(self):
        return (self._tag, self.kleft.spec(), self.kright.spec())

    def _emit(self, ctx, dims):
        el = self.kleft._emit(ctx, dims)
        er = self.kright._emit(ctx, dims)
        ctx["ops"].append([self._op, 0, 0, 0, 0, 0])
        return el + er


class SumKernel(_Pair):                  # sum_kernel.jl:1-16
    _op, _tag = OP["SUM"], "Sum"


class ProdKernel(_Pair):                 # prod_kernel.jl:1-15
    _op, _tag = OP["PROD"], "Prod"


class Masked(Kernel):                    # masked_kernel.jl:13-24 (active_dims 0-based here)
    def __init__(self, kernel, active_dims):
        self.kernel = kernel
        self.active_dims = [int(k) for k in active_dims]

    def get_params(self):
        return self.kernel.get_params()

    def set_params(self, hyp):
        self.kernel.set_params(hyp)

    def get_param_names(self):
        return self.kernel.get_param_names()

    def spec(self):
        return ("Masked", self.kernel.spec(), list(self.active_dims))

    def _emit(self, ctx, dims):
        return self.kernel._emit(ctx, [dims[k] for k in self.active_dims])


class FixedKernel(Kernel):               # fixed_kernel.jl:1-69 (free indices 0-based here)
    def __init__(self, kernel, free):
        self.kernel = kernel
        self.free = [int(f) for f in free]

    def get_params(self):
        p = self.kernel.get_params()
        return [p[f] for f in self.free]

    def set_params(self, hyp):
        if not self.free:
            return
        p = self.kernel.get_params()
        for f, v in zip(self.free, hyp):
            p[f] = float(v)
        self.kernel.set_params(p)

    def get_param_names(self):
        n = self.kernel.get_param_names()
        return [n[f] for f in self.free]

    def spec(self):
        return ("Fixed", self.kernel.spec(), list(self.free))

    def _emit(self, ctx, dims):
        e = self.kernel._emit(ctx, dims)
        return [e[f] for f in self.free]


def fix(k, par=None):                    # fixed_kernel.jl:24-45
    if isinstance(k, FixedKernel):
        names = k.kernel.get_param_names()
        return FixedKernel(k.kernel, [f for f in k.free if names[f] != par])
    if par is None:
        return FixedKernel(k, [])
    names = k.get_param_names()
    free = list(range(len(names)))
    if par in names:
        free.remove(names.index(par))
    return FixedKernel(k, free)


def flatten(kernel, d):
    """-> (ops int32[n_ops,6], dims int32[], theta_full float64[], exposed int[])

    `theta_full` is the parameter vector the device sees (all leaves, post-order == get_params order
    of the un-fixed tree); `exposed[i]` is the position in theta_full of the i-th entry of
    kernel.get_params() (FixedKernel hides the rest)."""
    ctx = dict(ops=[], dims=[], theta=[])
    exposed = kernel._emit(ctx, list(range(d)))
    if len(ctx["ops"]) > MAX_OPS or len(ctx["theta"]) > MAX_THETA or len(ctx["dims"]) > MAX_DIMS:
        raise ValueError("kernel program too large for the device interpreter")
    return (np.array(ctx["ops"], dtype=np.int32).reshape(-1, OP_STRIDE), np.array(ctx["dims"], dtype=np.int32),
            np.array(ctx["theta"], dtype=np.float64), exposed)
