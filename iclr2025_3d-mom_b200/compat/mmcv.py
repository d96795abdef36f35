"""Minimal `mmcv.Config.fromfile` (python config files with `_base_` inheritance), enough for
train_4DGS.py:440-443 / render_4DGS.py:110-113 and utils/params_utils.py:merge_hparams."""
import os


def _merge(base, new):
    out = dict(base)
    for k, v in new.items():
        if isinstance(v, dict) and isinstance(out.get(k), dict):
            out[k] = _merge(out[k], v)
        else:
            out[k] = v
    return out


def _load(path):
    scope = {}
    with open(path) as f:
        exec(compile(f.read(), path, "exec"), scope)
    cfg = {k: v for k, v in scope.items() if not k.startswith("__") and k != "_base_"}
    bases = scope.get("_base_", [])
    if isinstance(bases, str):
        bases = [bases]
    merged = {}
    for b in bases:
        merged = _merge(merged, _load(os.path.join(os.path.dirname(path), b)))
    return _merge(merged, cfg)


class Config(dict):
    @staticmethod
    def fromfile(filename):
        return Config(_load(os.path.abspath(filename)))

    def __getattr__(self, k):
        try:
            return self[k]
        except KeyError as e:
            raise AttributeError(k) from e
