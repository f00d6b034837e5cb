"""Serialise native builds across processes (pytest workers, torchrun ranks) and publish the result atomically."""
from __future__ import annotations

import contextlib
import fcntl
import os


@contextlib.contextmanager
def build_lock(target: str):
    """Exclusive lock for building `target`; yields a temporary output path that is renamed onto `target` on success."""
    lock_path = target + ".lock"
    tmp = f"{target}.tmp{os.getpid()}"
    with open(lock_path, "w") as lk:
        fcntl.flock(lk, fcntl.LOCK_EX)
        try:
            yield tmp
            if os.path.exists(tmp):
                os.replace(tmp, target)
        finally:
            if os.path.exists(tmp):
                os.unlink(tmp)
            fcntl.flock(lk, fcntl.LOCK_UN)
