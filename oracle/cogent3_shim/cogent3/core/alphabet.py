from cogent3 import AlphabetError  # noqa: F401
