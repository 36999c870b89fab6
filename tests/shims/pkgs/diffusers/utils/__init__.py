"""diffusers.utils: the handful of names the reference imports."""
import logging as _pylogging

USE_PEFT_BACKEND = True


class _Logging:
    @staticmethod
    def get_logger(name=None):
        return _pylogging.getLogger(name)


logging = _Logging()


def scale_lora_layers(model, weight):
    """peft.tuners.tuners_utils.scale_lora_layers: multiply every adapter's scaling by `weight` (no-op for 1.0)."""
    if weight == 1.0:
        return
    for m in model.modules():
        if hasattr(m, "scale_layer"):
            m.scale_layer(weight)


def unscale_lora_layers(model, weight=None):
    """peft.tuners.tuners_utils counterpart of scale_lora_layers: divide the scaling back (no-op for weight None / 1.0)."""
    if weight is None or weight == 1.0:
        return
    for m in model.modules():
        if hasattr(m, "unscale_layer"):
            m.unscale_layer(weight)


def convert_state_dict_to_diffusers(state_dict, original_type=None, **kwargs):
    """PEFT-format keys (`...lora_A.weight`) are already the diffusers LoRA format: identity for that input."""
    return state_dict


def is_torch_xla_available():
    """diffusers.utils.import_utils.is_torch_xla_available: no XLA here."""
    return False


def replace_example_docstring(example_docstring):
    """diffusers.utils.doc_utils.replace_example_docstring: decorator that splices an example into the docstring — a no-op for the tests."""
    return lambda fn: fn
