"""Mirror of the reference's `utils/` package for the accelerated path."""
from .ckpt import extract_model_state_dict, load_ckpt  # noqa: F401  (utils/__init__.py:4-26)
