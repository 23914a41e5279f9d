from . import environments  # noqa: F401
