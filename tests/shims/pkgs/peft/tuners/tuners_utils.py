from diffusers.utils import scale_lora_layers, unscale_lora_layers  # noqa: F401  (same restatement)
