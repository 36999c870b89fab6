from ..utils import USE_PEFT_BACKEND  # noqa: F401
from . import PeftAdapterMixin  # noqa: F401
