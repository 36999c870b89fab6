"""diffusers.configuration_utils: ConfigMixin + @register_to_config (init kwargs -> frozen `self.config`, attribute access)."""
import functools
import inspect


class FrozenDict(dict):
    """diffusers.configuration_utils.FrozenDict: the config mapping with attribute access (read-only in diffusers; reads are all the tests need)."""
    def __getattr__(self, k):
        try:
            return self[k]
        except KeyError as e:
            raise AttributeError(k) from e


class ConfigMixin:
    """diffusers.ConfigMixin: `self.config` = the __init__ arguments captured by @register_to_config."""
    config_name = "config.json"

    def register_to_config(self, **kwargs):
        """diffusers.configuration_utils.register_to_config: record the bound __init__ arguments (defaults included) in `self.config`."""
        kwargs.pop("kwargs", None)
        prev = dict(getattr(self, "_internal_dict", {}))
        self._internal_dict = FrozenDict({**prev, **kwargs})

    @property
    def config(self):
        return self._internal_dict


def register_to_config(init):
    """Decorator: every __init__ argument (given or default) except private `_x` ones is recorded in self.config."""

    @functools.wraps(init)
    def inner(self, *args, **kwargs):
        init_kwargs = {k: v for k, v in kwargs.items() if not k.startswith("_")}
        init(self, *args, **init_kwargs)
        sig = inspect.signature(init)
        params = {n: p.default for i, (n, p) in enumerate(sig.parameters.items()) if i > 0}
        new = {}
        for a, name in zip(args, params.keys()):
            new[name] = a
        new.update({k: init_kwargs.get(k, d) for k, d in params.items() if k not in new})
        getattr(self, "register_to_config")(**new)

    return inner
