from . import perspective  # noqa: F401
