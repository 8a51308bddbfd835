"""``models.trimap.model`` of the reference (models/trimap/model.py:173): the eval wrapper, B200-backed."""
from otvm_b200.models import FullModel_eval  # noqa: F401
