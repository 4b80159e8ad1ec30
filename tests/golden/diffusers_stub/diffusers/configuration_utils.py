"""register_to_config of diffusers.configuration_utils, reduced to what the reference reads
(`self.config.<ctor kwarg>`): the bound constructor arguments (defaults included) of every
decorated __init__ in the MRO are merged into one dict."""
import functools
import inspect


def register_to_config(init):
    @functools.wraps(init)
    def inner(self, *args, **kwargs):
        sig = inspect.signature(init)
        names = [n for i, (n, p) in enumerate(sig.parameters.items())
                 if i > 0 and p.kind not in (p.VAR_KEYWORD, p.VAR_POSITIONAL)]
        cfg = {n: sig.parameters[n].default for n in names}
        cfg.update(dict(zip(names, args)))
        cfg.update({k: v for k, v in kwargs.items() if not k.startswith("_")})
        merged = dict(self.__dict__.get("_config_dict", {}))
        merged.update(cfg)
        self.__dict__["_config_dict"] = merged
        init(self, *args, **{k: v for k, v in kwargs.items() if not k.startswith("_")})
    return inner
