"""Auto-stub importer (tests only).  The reference's trainer modules import many names they never touch on the hot path
(pipelines, encoders, tokenizers, accelerate, omegaconf, ...).  After `install()`, any module under the whitelisted prefixes that
neither the real environment nor tests/shims provides resolves to an empty stub whose attributes are inert placeholder classes —
so `import qflux.trainer.qwen_image_edit_trainer` succeeds and the methods on the §8 path (`_compute_loss`, `clip_gradients`,
`add_lora_adapter`, `save_lora`, ...) can be driven against lightweight `self` objects.  Anything SEMANTIC (diffusers leaf
layers, peft LoRA, scheduler tables, density / weighting helpers) is restated explicitly in tests/shims/{diffusers,peft,accelerate}."""
import importlib.abc
import importlib.machinery
import importlib.util
import os
import sys
import types

PKGS = os.path.join(os.path.dirname(os.path.abspath(__file__)), "pkgs")
REAL = ("transformers", "wandb")  # installed for real: only their MISSING submodules are stubbed

PREFIXES = ("diffusers", "peft", "accelerate", "omegaconf", "imagehash", "bitsandbytes", "prodigyopt", "transformers", "optimum", "lpips", "swanlab", "tensorboardX")


class _StubMeta(type):
    def __getattr__(cls, name):  # class-level attribute access (e.g. `Pipeline._pack_latents`)
        if name.startswith("__"):
            raise AttributeError(name)
        return _Stub()


class _Stub(metaclass=_StubMeta):
    """inert placeholder: constructible, callable, attribute access yields more placeholders"""

    def __init__(self, *a, **k):
        pass

    def __call__(self, *a, **k):
        return _Stub()

    def __getattr__(self, name):
        if name.startswith("__"):
            raise AttributeError(name)
        return _Stub()


class _StubModule(types.ModuleType):
    def __getattr__(self, name):
        if name.startswith("__"):
            raise AttributeError(name)
        cls = type(name, (_Stub,), {"__module__": self.__name__})
        setattr(self, name, cls)
        return cls


class _ShimLoader(importlib.machinery.SourceFileLoader):
    """runs a restatement file, then lets names it does not define resolve to inert placeholders (PEP 562 module __getattr__)"""

    def exec_module(self, module):
        super().exec_module(module)
        if hasattr(module, "__path__"):
            # sub-modules must come back through _Finder (importlib re-populates an empty search path with the directory, and the default
            # path finder would then load them WITHOUT the placeholder fallback below: `from pkg.sub import Missing` would yield a stub
            # module instead of a class)
            module.__path__ = []
        own = module.__dict__.get("__getattr__")

        def fallback(name, _own=own, _mod=module):
            if _own is not None:
                try:
                    return _own(name)
                except AttributeError:
                    pass
            if name.startswith("__"):
                raise AttributeError(name)
            cls = type(name, (_Stub,), {"__module__": _mod.__name__})
            setattr(_mod, name, cls)
            return cls
        module.__getattr__ = fallback


class _Finder(importlib.abc.MetaPathFinder, importlib.abc.Loader):
    def find_spec(self, fullname, path=None, target=None):
        # availability probes (`importlib.util.find_spec("accelerate")`, as transformers / torch do) must keep seeing "not installed":
        # only a real `import` statement gets a stub
        f = sys._getframe(1)
        for _ in range(8):
            if f is None:
                break
            if f.f_code.co_name == "find_spec" and f.f_code.co_filename.replace("\\", "/").endswith("importlib/util.py"):
                return None
            f = f.f_back
        if not any(fullname == p or fullname.startswith(p + ".") for p in PREFIXES):
            return None
        # semantic restatements live under tests/shims/pkgs (NOT on sys.path, so that probes never see them)
        base = os.path.join(PKGS, *fullname.split("."))
        if os.path.isdir(base) and os.path.exists(os.path.join(base, "__init__.py")):
            f = os.path.join(base, "__init__.py")
            return importlib.util.spec_from_file_location(fullname, f, loader=_ShimLoader(fullname, f), submodule_search_locations=[])
        if os.path.exists(base + ".py"):
            return importlib.util.spec_from_file_location(fullname, base + ".py", loader=_ShimLoader(fullname, base + ".py"))
        if fullname.split(".")[0] in REAL and path is None:
            return None
        return importlib.machinery.ModuleSpec(fullname, self, is_package=True)

    def create_module(self, spec):
        m = _StubModule(spec.name)
        m.__path__ = []
        return m

    def exec_module(self, module):
        pass


_installed = None


def install():
    """Append the finder: really installed modules always win; tests/shims/pkgs next; inert stubs last."""
    global _installed
    if _installed is None:
        import transformers  # noqa: F401  (its cached availability probes must run BEFORE any stand-in sits in sys.modules)
        _installed = _Finder()
        sys.meta_path.append(_installed)
    os.environ.setdefault("QFLUX_DOTENV_LOADED", "1")  # skips the .env / HF-login side effect of qflux/__init__.py:12-17
    return _installed
