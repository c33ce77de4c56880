"""Message / exit helpers with the reference's behaviour (NFACT/base/utils.py:8-123)."""
import logging
import re
import sys
import time


class Timer:
    """NFACT/base/utils.py:8-38."""

    def __init__(self):
        self._t = None

    def tic(self):
        self._t = time.time()

    def toc(self):
        return f"{time.time() - self._t:.2f}"

    def how_long(self):
        seconds = int(round(float(self.toc())))
        if seconds < 60:
            return f"{seconds} seconds"
        minutes, rem = divmod(seconds, 60)
        if seconds < 3600:
            return f"{minutes} minutes: {rem} seconds"
        hours, minutes = divmod(minutes, 60)
        return f"{hours} hours: {minutes} minutes: {rem} seconds"


def colours() -> dict:
    return {"reset": "\033[0;0m", "red": "\033[1;31m", "pink": "\033[1;35m"}


def nprint(message: str, to_flush: bool = False) -> None:
    """Print and try to log (NFACT/base/utils.py:101-123)."""
    print(message, flush=to_flush)
    try:
        logging.info(message)
    except Exception:
        return None


def error_and_exit(bool_statement: bool, error_message: str = None, log_message: bool = True) -> None:
    """Red message then ``exit(1)`` when the statement is false (NFACT/base/utils.py:66-98)."""
    if not bool_statement:
        if error_message:
            error_message = re.sub(r"\[Errno 2\]", "", error_message)
            col = colours()
            message = col["red"] + error_message + col["reset"]
            nprint(message) if log_message else print(message)
        print("Exiting...\n")
        sys.exit(1)
