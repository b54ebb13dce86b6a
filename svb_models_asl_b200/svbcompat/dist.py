"""
Distributions and internal<->external value transforms for model parameters.

Host-side mirror of the external ``svb.dist`` module that the reference plugins
import (``/root/reference/svb_models_asl/aslnn.py:28``) and name through
``get_parameter(dist="Normal"|"LogNormal"|"FoldedNormal")``
(``aslrest.py:184``, ``aslrest_disp.py:34,37``, ``aslnn.py:74,78``).  The svb
source is not part of the reference tree, so the semantics here follow
SURVEY.md Appendix B: the posterior is always Gaussian over an *internal*
value theta, and the model sees ``ext = transform(theta)``.
"""
import math

import numpy as np

# Transform codes shared with the CUDA side (include/svbasl.h: SVBASL_XF_*)
XF_IDENTITY, XF_EXP, XF_ABS = 0, 1, 2


class _Transform:
    code = XF_IDENTITY

    def int_values(self, ext):
        return ext

    def ext_values(self, internal):
        return internal


class Identity(_Transform):
    pass


class Log(_Transform):
    """Model sees exp(theta); internal value is the log of the model value."""
    code = XF_EXP

    def int_values(self, ext):
        return np.log(ext)

    def ext_values(self, internal):
        return np.exp(internal)


class Abs(_Transform):
    """Model sees |theta|."""
    code = XF_ABS

    def ext_values(self, internal):
        return np.abs(internal)


class Dist:
    pass


class Normal(Dist):
    def __init__(self, mean, var, **_kw):
        self.transform = Identity()
        self.mean, self.var = mean, var
        self.sd = np.sqrt(var)

    def __str__(self):
        return "Gaussian (%s, %s)" % (_fmt(self.mean), _fmt(self.var))


class LogNormal(Normal):
    """
    Log of the value is Gaussian.  With ``geom`` (the default) the supplied
    mean/variance are taken as geometric moments, so the internal Gaussian is
    N(log mean, log var); otherwise they are moment-matched.
    """
    def __init__(self, mean, var, geom=True, **kw):
        self.ext_mean, self.ext_var = mean, var
        if geom:
            nmean, nvar = np.log(mean), np.log(var)
        else:
            nmean = np.log(mean ** 2 / np.sqrt(mean ** 2 + var))
            nvar = np.log(mean ** 2 / var + 1)
        Normal.__init__(self, nmean, nvar, **kw)
        self.transform = Log()

    def __str__(self):
        return "Log-Normal (%s, %s)" % (_fmt(self.ext_mean), _fmt(self.ext_var))


class FoldedNormal(Normal):
    def __init__(self, mean, var, **kw):
        Normal.__init__(self, mean, var, **kw)
        self.transform = Abs()

    def __str__(self):
        return "Folded Normal (%s, %s)" % (_fmt(self.mean), _fmt(self.var))


def _fmt(v):
    a = np.asarray(v)
    return "%g" % float(a) if a.ndim == 0 else "array[%i]" % a.size


_KNOWN = {"Normal": Normal, "LogNormal": LogNormal, "FoldedNormal": FoldedNormal}


def get_dist(prefix, **kwargs):
    """``<prefix>_dist|dist``, ``<prefix>_mean|mean``, ``<prefix>_var|var`` -> Dist"""
    name = kwargs.get("%s_dist" % prefix, kwargs.get("dist", "Normal"))
    mean = kwargs.get("%s_mean" % prefix, kwargs.get("mean", 0.0))
    var = kwargs.get("%s_var" % prefix, kwargs.get("var", 1.0))
    if name not in _KNOWN:
        raise ValueError("Unrecognized distribution: %s" % name)
    return _KNOWN[name](mean, var)
