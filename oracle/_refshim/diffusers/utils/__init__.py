from dataclasses import fields, is_dataclass
from collections import OrderedDict
import logging as _pylogging

import torch


class BaseOutput(OrderedDict):
    """Dataclass-style output that also supports tuple/dict access (diffusers.utils.BaseOutput)."""

    def __init_subclass__(cls):
        super().__init_subclass__()

    def __post_init__(self):
        for f in fields(self):
            v = getattr(self, f.name)
            if v is not None:
                self[f.name] = v

    def __getitem__(self, k):
        if isinstance(k, str):
            return dict(self.items())[k]
        return self.to_tuple()[k]

    def to_tuple(self):
        return tuple(self[k] for k in self.keys())


def is_torch_version(op, ver):
    from packaging import version
    import operator
    ops = {">=": operator.ge, ">": operator.gt, "<": operator.lt, "<=": operator.le, "==": operator.eq}
    return ops[op](version.parse(torch.__version__.split("+")[0]), version.parse(ver))


class logging:  # noqa: N801
    @staticmethod
    def get_logger(name):
        return _pylogging.getLogger(name)
