"""
Model parameter description (prior, initial posterior, prior type, initialiser).

Mirror of ``svb.parameter.get_parameter`` as called by the reference plugins
(``/root/reference/svb_models_asl/aslrest.py:13,184-246``).  Keyword rules:
``mean/var`` apply to prior and posterior unless ``prior_mean/prior_var`` or
``post_mean/post_var`` override them; ``prior_type`` is "N" (fixed Normal),
"A" (ARD, ``aslrest.py:237``) or "M" (spatial MRF); the whole option dict of
the model is forwarded (``**options``) and may carry ``param_overrides``
``{name: {...}}`` which win over everything else.
"""
from . import dist as _dist


class Parameter:
    def __init__(self, name, prior, post, prior_type="N", post_init=None, desc=""):
        self.name = name
        self.desc = desc
        self.prior_dist = prior
        self.post_dist = post
        self.prior_type = prior_type
        self.post_init = post_init

    def __str__(self):
        return "Parameter: %s" % self.name


def get_parameter(name, **kwargs):
    kwargs = dict(kwargs)
    overrides = (kwargs.pop("param_overrides", None) or {}).get(name, {})
    kwargs.update(overrides)
    return Parameter(
        name,
        prior=_dist.get_dist("prior", **kwargs),
        post=_dist.get_dist("post", **kwargs),
        prior_type=kwargs.get("prior_type", "N"),
        post_init=kwargs.get("post_init", None),
        desc=kwargs.get("desc", "No description given"),
    )
