"""Error taxonomy and enforce helpers. Parity: paddle/common/errors.h (error codes), paddle/common/enforce.h (PADDLE_ENFORCE_*,
PADDLE_THROW), python/paddle/base/core EnforceNotMet.  Every error type derives from EnforceNotMet AND from the closest Python
built-in, so `except ValueError` style user code keeps working."""
from __future__ import annotations

import traceback

__all__ = ["EnforceNotMet", "InvalidArgumentError", "NotFoundError", "OutOfRangeError", "AlreadyExistsError", "ResourceExhaustedError",
           "PreconditionNotMetError", "PermissionDeniedError", "ExecutionTimeoutError", "UnimplementedError", "UnavailableError",
           "FatalError", "ExternalError", "InvalidTypeError", "ErrorCode", "enforce", "enforce_eq", "enforce_ne", "enforce_gt", "enforce_ge",
           "enforce_lt", "enforce_le", "enforce_not_none", "enforce_shape_match", "throw"]


class ErrorCode:
    LEGACY, INVALID_ARGUMENT, NOT_FOUND, OUT_OF_RANGE, ALREADY_EXISTS, RESOURCE_EXHAUSTED, PRECONDITION_NOT_MET, PERMISSION_DENIED, \
        EXECUTION_TIMEOUT, UNIMPLEMENTED, UNAVAILABLE, FATAL, EXTERNAL, INVALID_TYPE = range(14)


class EnforceNotMet(RuntimeError):
    code = ErrorCode.LEGACY
    kind = "Error"

    def __init__(self, message="", hint=None):
        self.raw_message = str(message)
        self.hint = hint
        text = f"({self.kind}) {self.raw_message}"
        if hint:
            text += f"\n  [Hint: {hint}]"
        super().__init__(text)


def _mk(name, code, kind, *bases):
    return type(name, (EnforceNotMet,) + bases, {"code": code, "kind": kind, "__doc__": f"{kind} error (code {code})."})


InvalidArgumentError = _mk("InvalidArgumentError", ErrorCode.INVALID_ARGUMENT, "InvalidArgument", ValueError)
NotFoundError = _mk("NotFoundError", ErrorCode.NOT_FOUND, "NotFound", LookupError)
OutOfRangeError = _mk("OutOfRangeError", ErrorCode.OUT_OF_RANGE, "OutOfRange", IndexError)
AlreadyExistsError = _mk("AlreadyExistsError", ErrorCode.ALREADY_EXISTS, "AlreadyExists")
ResourceExhaustedError = _mk("ResourceExhaustedError", ErrorCode.RESOURCE_EXHAUSTED, "ResourceExhausted", MemoryError)
PreconditionNotMetError = _mk("PreconditionNotMetError", ErrorCode.PRECONDITION_NOT_MET, "PreconditionNotMet")
PermissionDeniedError = _mk("PermissionDeniedError", ErrorCode.PERMISSION_DENIED, "PermissionDenied", PermissionError)
ExecutionTimeoutError = _mk("ExecutionTimeoutError", ErrorCode.EXECUTION_TIMEOUT, "ExecutionTimeout", TimeoutError)
UnimplementedError = _mk("UnimplementedError", ErrorCode.UNIMPLEMENTED, "Unimplemented", NotImplementedError)
UnavailableError = _mk("UnavailableError", ErrorCode.UNAVAILABLE, "Unavailable")
FatalError = _mk("FatalError", ErrorCode.FATAL, "Fatal")
ExternalError = _mk("ExternalError", ErrorCode.EXTERNAL, "External", OSError)
InvalidTypeError = _mk("InvalidTypeError", ErrorCode.INVALID_TYPE, "InvalidType", TypeError)


def throw(error_type, message, *args):
    """PADDLE_THROW: raise `error_type` with a printf-style message."""
    raise error_type(message % args if args else message)


def enforce(cond, error_type=InvalidArgumentError, message="enforce failed", *args):
    """PADDLE_ENFORCE: raise unless `cond` holds. `cond` may be a bool or a 0-d tensor (read back once)."""
    ok = bool(cond.item()) if hasattr(cond, "item") and not isinstance(cond, bool) else bool(cond)
    if not ok:
        caller = traceback.extract_stack(limit=2)[0]
        raise error_type(message % args if args else message, hint=f"at {caller.filename}:{caller.lineno}")


def _cmp(name, op):
    def f(a, b, error_type=InvalidArgumentError, message=None):
        if not op(a, b):
            raise error_type(message or f"Expected {a!r} {name} {b!r}.")
    f.__name__ = f"enforce_{name}"
    return f


enforce_eq = _cmp("==", lambda a, b: a == b)
enforce_ne = _cmp("!=", lambda a, b: a != b)
enforce_gt = _cmp(">", lambda a, b: a > b)
enforce_ge = _cmp(">=", lambda a, b: a >= b)
enforce_lt = _cmp("<", lambda a, b: a < b)
enforce_le = _cmp("<=", lambda a, b: a <= b)


def enforce_not_none(v, what="value", error_type=NotFoundError):
    if v is None:
        raise error_type(f"{what} should not be null.")
    return v


def enforce_shape_match(a, b, what="shapes"):
    if list(a) != list(b):
        raise InvalidArgumentError(f"{what} mismatch: {list(a)} vs {list(b)}.")
