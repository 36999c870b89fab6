"""diffusers.loaders: PeftAdapterMixin (add_adapter / set_adapter / load_lora_adapter via peft injection) and empty mixins."""
import torch.nn as nn


class FromOriginalModelMixin:
    """diffusers FromOriginalModelMixin (single-file checkpoint loading): not used by the tests, present as a base class."""
    pass


class FluxTransformer2DLoadersMixin:
    """diffusers FluxTransformer2DLoadersMixin (IP-adapter weight loading): not used by the tests, present as a base class."""
    pass


class AttnProcsLayers(nn.Module):
    """diffusers.loaders.AttnProcsLayers: an nn.Module holding the given {name: module} dict as a ModuleList — what the reference wraps
    `get_lora_layers(dit)` in before `accelerator.prepare` (base_trainer.py:384)."""
    def __init__(self, state_dict):
        super().__init__()
        self.layers = nn.ModuleList(state_dict.values())


class PeftAdapterMixin:
    """diffusers.loaders.peft.PeftAdapterMixin (the three entry points the reference trainer calls, base_trainer.py:940-941,983)."""
    _hf_peft_config_loaded = False

    def add_adapter(self, adapter_config, adapter_name: str = "default"):
        from peft import inject_adapter_in_model
        if getattr(self, "peft_config", None) and adapter_name in self.peft_config:
            raise ValueError(f"Adapter with name {adapter_name} already exists. Please use a different name.")
        inject_adapter_in_model(adapter_config, self, adapter_name)
        self._hf_peft_config_loaded = True
        self.set_adapter(adapter_name)

    def set_adapter(self, adapter_name):
        names = [adapter_name] if isinstance(adapter_name, str) else list(adapter_name)
        for m in self.modules():
            if hasattr(m, "lora_A") and hasattr(m, "set_adapter") and m is not self:
                m.set_adapter(names)

    def load_lora_adapter(self, pretrained_model_name_or_path_or_dict, prefix="transformer", adapter_name="default", **kwargs):
        """Reads a diffusers-format LoRA file (keys `transformer.<module>.lora_A.weight`), infers r per module from lora_B,
        alpha = r (no network_alphas in this format), injects and loads."""
        import safetensors.torch
        from peft import LoraConfig, inject_adapter_in_model, set_peft_model_state_dict
        sd = pretrained_model_name_or_path_or_dict
        if not isinstance(sd, dict):
            sd = safetensors.torch.load_file(sd)
        if prefix is not None and any(k.startswith(prefix + ".") for k in sd):
            sd = {k[len(prefix) + 1:]: v for k, v in sd.items() if k.startswith(prefix + ".")}
        ranks = {k.rsplit(".lora_B", 1)[0]: v.shape[1] for k, v in sd.items() if ".lora_B" in k and v.ndim > 1}
        r = max(set(ranks.values()), key=list(ranks.values()).count)
        cfg = LoraConfig(r=r, lora_alpha=r, target_modules=sorted(ranks), init_lora_weights=True,
                         rank_pattern={k: v for k, v in ranks.items() if v != r})
        inject_adapter_in_model(cfg, self, adapter_name)
        set_peft_model_state_dict(self, sd, adapter_name)
        self._hf_peft_config_loaded = True
        self.set_adapter(adapter_name)
