import contextlib


class CacheMixin:
    """diffusers.models.cache_utils.CacheMixin: inference-time caching hooks; no effect on the training forward."""

    def cache_context(self, name: str):
        """`with transformer.cache_context("cond"): ...` names the active branch for the cache hooks; without a cache enabled (the only
        case here) diffusers' implementation yields without doing anything."""
        return contextlib.nullcontext()
