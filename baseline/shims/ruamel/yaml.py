"""Minimal `ruamel.yaml` stand-in (harness only, NOT product code).

The reference's template loader (pyrates/frontend/fileio/yaml.py:69-74) only needs
``YAML(typ="safe", pure=True).load(file)``.  PyYAML implements YAML 1.1, whose float
resolver rejects exponent literals without a dot (``10e-3``), which the reference's
model templates use; ruamel implements YAML 1.2 where these are floats.  This shim
adds a YAML-1.2-style float resolver on top of PyYAML's SafeLoader.
"""
import re
import yaml as _pyyaml


class _Loader12(_pyyaml.SafeLoader):
    pass


# drop PyYAML's 1.1 float resolver, install a 1.2 one
_Loader12.yaml_implicit_resolvers = {
    k: [(tag, rx) for tag, rx in v if tag != 'tag:yaml.org,2002:float']
    for k, v in _pyyaml.SafeLoader.yaml_implicit_resolvers.items()
}
_Loader12.add_implicit_resolver(
    'tag:yaml.org,2002:float',
    re.compile(r'''^(?:[-+]?(?:[0-9][0-9_]*)\.[0-9_]*(?:[eE][-+]?[0-9]+)?
                    |[-+]?(?:[0-9][0-9_]*)(?:[eE][-+]?[0-9]+)
                    |[-+]?\.[0-9_]+(?:[eE][-+]?[0-9]+)?
                    |[-+]?\.(?:inf|Inf|INF)
                    |\.(?:nan|NaN|NAN))$''', re.X),
    list('-+0123456789.'))


class YAML:
    def __init__(self, typ="safe", pure=True, **_):
        self.typ = typ

    def load(self, stream):
        return _pyyaml.load(stream, Loader=_Loader12)

    def dump(self, data, stream=None, **kw):
        return _pyyaml.safe_dump(data, stream, **kw)
