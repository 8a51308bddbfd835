"""otvm_b200 - Blackwell (sm_100a) implementation of the OTVM per-frame inference hot path."""
from .helpers import get_model_alpha, get_model_name, get_model_trimap  # noqa: F401

__all__ = ["get_model_trimap", "get_model_alpha", "get_model_name"]
