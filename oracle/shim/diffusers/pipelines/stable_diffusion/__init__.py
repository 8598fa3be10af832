from . import safety_checker  # noqa: F401
