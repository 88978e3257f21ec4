"""Import stub (test infrastructure) for the reference's `from gymnasium import spaces`."""
from . import spaces  # noqa: F401
