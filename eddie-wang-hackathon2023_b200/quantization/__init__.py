from .mode import QuantMode  # noqa: F401
from . import functional  # noqa: F401
from . import layer  # noqa: F401
