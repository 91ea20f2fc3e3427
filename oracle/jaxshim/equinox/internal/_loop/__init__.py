from . import checkpointed  # noqa: F401
