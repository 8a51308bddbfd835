"""``models.alpha.model`` of the reference (models/alpha/model.py:314): the eval model, B200-backed."""
from otvm_b200.models import EvalModel  # noqa: F401
