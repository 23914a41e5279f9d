from . import distributed  # noqa: F401
