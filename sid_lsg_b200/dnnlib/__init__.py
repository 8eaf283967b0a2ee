"""The three `dnnlib` names the loop uses (/root/reference/dnnlib/util.py): `EasyDict`,
`util.construct_class_by_name`, `util.format_time`.  Fresh implementations."""
from . import util  # noqa: F401
from .util import EasyDict  # noqa: F401
