from . import experimental  # noqa: F401
