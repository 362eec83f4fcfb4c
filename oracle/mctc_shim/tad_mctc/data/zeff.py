from ._tables import ZEFF  # noqa: F401
