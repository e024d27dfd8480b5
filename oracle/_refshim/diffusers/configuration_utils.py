"""ConfigMixin / register_to_config: the subset used at
hyvideo/vae/autoencoder_kl_causal_3d.py:25,63 and hyvideo/vae/__init__.py:88-92."""
import functools
import inspect
import json
import os


class FrozenDict(dict):
    def __getattr__(self, k):
        try:
            return self[k]
        except KeyError as e:
            raise AttributeError(k) from e


def register_to_config(init):
    @functools.wraps(init)
    def inner(self, *args, **kwargs):
        sig = inspect.signature(init)
        cfg = {n: p.default for n, p in list(sig.parameters.items())[1:] if p.default is not inspect._empty}
        names = [n for n in list(sig.parameters)[1:]]
        for n, a in zip(names, args):
            cfg[n] = a
        cfg.update(kwargs)
        self._internal_dict = FrozenDict(cfg)
        init(self, *args, **kwargs)
    return inner


class ConfigMixin:
    config_name = "config.json"

    @property
    def config(self):
        return self._internal_dict

    @classmethod
    def load_config(cls, path, **kw):
        p = path if str(path).endswith(".json") else os.path.join(path, cls.config_name)
        with open(p) as f:
            return json.load(f)

    @classmethod
    def from_config(cls, config, **kwargs):
        sig = inspect.signature(cls.__init__)
        cfg = {k: v for k, v in dict(config).items() if k in sig.parameters}
        cfg.update(kwargs)
        return cls(**cfg)
