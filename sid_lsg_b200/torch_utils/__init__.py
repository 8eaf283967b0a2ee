"""Host-side mirror of the reference's `torch_utils` surface that the distillation loop touches
(/root/reference/torch_utils/{distributed,misc,training_stats}.py).  Written fresh: same names, argument meaning
and error behaviour; none of the reference's bodies."""
from . import distributed, misc, training_stats  # noqa: F401
