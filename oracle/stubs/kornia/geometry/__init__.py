from . import camera  # noqa: F401
