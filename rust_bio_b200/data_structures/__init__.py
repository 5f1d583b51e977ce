from .bitenc import BitEnc  # noqa: F401
