"""Error types raised by the host layer.  Inside the reference process the reference's own
`GeneralError` (gprMax/exceptions.py:29-36) is used so callers that catch it keep working."""
try:  # running as a drop-in inside gprMax
    from gprMax.exceptions import GeneralError  # noqa: F401
except Exception:  # standalone
    class GeneralError(ValueError):
        """Handles general errors. Subclasses the ValueError class (gprMax/exceptions.py:29)."""

        def __init__(self, message, *args):
            self.message = message
            super(GeneralError, self).__init__(message, *args)
