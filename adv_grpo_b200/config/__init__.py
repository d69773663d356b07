"""Config surface of the reference (`config/base.py`, `config/grpo.py`, `--config file.py:<preset>`):
an `ml_collections.ConfigDict`-compatible container (the package is not installed here), the presets
the two hot-path scripts use, and a Python-3.12-safe loader that can also execute the REFERENCE's own
config files (they `import imp`, removed in 3.12, and `import ml_collections`)."""
import importlib.util
import os
import sys
import types


class ConfigDict(dict):
    """Attribute + item access, nested dict promotion, `.get`, `.to_dict()` -- the subset of
    ml_collections.ConfigDict that config/*.py and the training scripts touch."""

    def __init__(self, initial=None):
        super().__init__()
        for k, v in (initial or {}).items():
            self[k] = v

    def __setitem__(self, k, v):
        if isinstance(v, dict) and not isinstance(v, ConfigDict):
            v = ConfigDict(v)
        super().__setitem__(k, v)

    def __getattr__(self, k):
        try:
            return self[k]
        except KeyError:
            raise AttributeError(k)

    def __setattr__(self, k, v):
        self[k] = v

    def to_dict(self):
        return {k: (v.to_dict() if isinstance(v, ConfigDict) else v) for k, v in self.items()}

    def items(self):
        return super().items()


def install_shims():
    """Make `import ml_collections` and `import imp` work for reference config files."""
    if "ml_collections" not in sys.modules:
        m = types.ModuleType("ml_collections")
        m.ConfigDict = ConfigDict
        m.__advgrpo_shim__ = True
        sys.modules["ml_collections"] = m
    if "imp" not in sys.modules:
        imp = types.ModuleType("imp")

        def load_source(name, path):
            spec = importlib.util.spec_from_file_location(name, path)
            mod = importlib.util.module_from_spec(spec)
            spec.loader.exec_module(mod)
            return mod

        imp.load_source = load_source
        sys.modules["imp"] = imp


def load_config(spec):
    """`path/to/grpo.py:preset` (config_flags syntax) or a preset name of this package."""
    install_shims()
    if ":" in spec or spec.endswith(".py"):
        path, _, name = spec.partition(":")
        spec_ = importlib.util.spec_from_file_location("advgrpo_user_config", os.path.abspath(path))
        mod = importlib.util.module_from_spec(spec_)
        spec_.loader.exec_module(mod)
        return mod.get_config(name) if name else mod.get_config()
    from . import grpo
    return grpo.get_config(spec)
