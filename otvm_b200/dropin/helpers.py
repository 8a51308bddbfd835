"""Drop-in replacement for the reference's top-level ``helpers`` module (helpers.py:323-362).

    PYTHONPATH=/path/to/otvm_b200/dropin:/path/to/repo python eval.py --gpu 0

``eval.py:74-75`` calls ``helpers.get_model_trimap`` / ``helpers.get_model_alpha``; with this directory first on the
path those return the B200 implementation, everything else in eval.py runs unchanged."""
from otvm_b200.helpers import get_model_alpha, get_model_name, get_model_trimap  # noqa: F401
