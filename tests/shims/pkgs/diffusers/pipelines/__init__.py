"""The one pipeline classmethod the trainer's save path uses (base_trainer.py:858-876): `save_lora_weights`."""
import os


class _AnyClassAttr(type):
    def __getattr__(cls, name):  # pipeline helpers the trainers alias at class-definition time but the hot path never calls
        if name.startswith("__"):
            raise AttributeError(name)

        def _unavailable(*a, **k):
            raise NotImplementedError(f"{cls.__name__}.{name} is not restated in tests/shims")
        return _unavailable


class LoraSavingPipeline(metaclass=_AnyClassAttr):
    """Stands in for every `*Pipeline` class the trainers name as `pipeline_class`: `save_lora_weights(dir, transformer_lora_layers, ...)`
    writes `pytorch_lora_weights.safetensors` with the `transformer.` key prefix, like diffusers' LoraBaseMixin.write_lora_layers."""
    transformer_name = "transformer"

    @classmethod
    def save_lora_weights(cls, save_directory, transformer_lora_layers=None, safe_serialization=True, weight_name=None, **kwargs):
        """diffusers LoraBaseMixin.write_lora_layers: keys prefixed with `transformer.`, file pytorch_lora_weights.safetensors."""
        import safetensors.torch
        os.makedirs(save_directory, exist_ok=True)
        sd = {f"{cls.transformer_name}.{k}": v.detach().contiguous() for k, v in transformer_lora_layers.items()}
        path = os.path.join(save_directory, weight_name or "pytorch_lora_weights.safetensors")
        safetensors.torch.save_file(sd, path, metadata={"format": "pt"})
        return path
