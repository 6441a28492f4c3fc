"""Exceptions of the augment path.

The reference has no error type of its own: malformed input surfaces as an
uncaught IndexError / KeyError / ValueError / AssertionError, i.e. a traceback on
stderr, a non-zero exit status and nothing on stdout.
The replacement keeps that contract with two named exceptions.
"""


class PantasError(RuntimeError):
    pass


class PantasDataError(PantasError):
    """A record on which the reference script itself raises."""

    def __init__(self, msg, code=None, offset=None):
        super().__init__(msg)
        self.code = code
        self.offset = offset


class UnsupportedInput(PantasError):
    """Input the reference would accept by relying on Python behaviour that the
    device parser refuses to guess at (documented deviations)."""

    def __init__(self, msg, code=None, offset=None):
        super().__init__(msg)
        self.code = code
        self.offset = offset


class NativeLibraryError(PantasError):
    """libpantas_aug.so is missing, not loadable, or reported an API error.
    There is no CPU fallback: the product path stops here."""
