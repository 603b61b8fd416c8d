"""Stand-in for `omegaconf` (absent offline): OmegaConf.load / merge / save over PyYAML, returning attribute dicts.
Only what /root/reference/main.py uses (a 3-level merge of YAML files)."""
import yaml
from easydict import EasyDict


def _merge(a, b):
    out = dict(a)
    for k, v in b.items():
        out[k] = _merge(out[k], v) if isinstance(v, dict) and isinstance(out.get(k), dict) else v
    return out


def _plain(d):
    if isinstance(d, dict):
        return {k: _plain(v) for k, v in d.items()}
    if isinstance(d, (list, tuple)):
        return [_plain(v) for v in d]
    return d


class OmegaConf:
    @staticmethod
    def load(path):
        with open(path) as f:
            return EasyDict(yaml.safe_load(f) or {})

    @staticmethod
    def merge(*cfgs):
        out = {}
        for c in cfgs:
            out = _merge(out, _plain(c))
        return EasyDict(out)

    @staticmethod
    def create(d=None):
        return EasyDict(d or {})

    @staticmethod
    def save(cfg, f):
        yaml.safe_dump(_plain(cfg), f)
