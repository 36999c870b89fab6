class CacheMixin:
    """diffusers.models.cache_utils.CacheMixin: inference-time caching hooks; no effect on the training forward."""
