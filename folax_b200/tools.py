"""Small host utilities with the reference's behaviour (fol/tools/decoration_functions.py:34-116)."""
import warnings
from datetime import datetime


def _stamp():
    return datetime.now().strftime("%Y-%m-%d %H:%M:%S")


def fol_info(message, who=""):
    print(f"{_stamp()} - Info : {who}{' - ' if who else ''}{message}")


def fol_warning(message, who=""):
    warnings.warn(f"{_stamp()} - Warning : {who}{' - ' if who else ''}{message}", UserWarning, stacklevel=2)


def fol_error(message, who=""):
    """Prints and stops execution with SystemExit, as the reference does (:60-87)."""
    print(f"{_stamp()} - Error : {who}{' - ' if who else ''}{message}")
    raise SystemExit
