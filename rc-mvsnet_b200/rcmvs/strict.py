"""Library-fallback accounting.

Every place where a drop-in module would hand work to a torch library kernel (cuDNN / cuBLAS / ATen) instead of a
librcmvs kernel calls `library_fallback(site)`:
    RCMVS_STRICT=1  -> RuntimeError (tests, bench.py and smoke() run like this: a silent fallback cannot pass for ours)
    otherwise       -> one warning per site per process, then the torch path runs.
"""
import os
import warnings

_seen = set()


def strict():
    return os.environ.get("RCMVS_STRICT", "0") == "1"


def library_fallback(site, why=""):
    msg = "rcmvs: %s runs on torch library kernels, not on librcmvs%s" % (site, (" (%s)" % why) if why else "")
    if strict():
        raise RuntimeError(msg + " [RCMVS_STRICT=1]")
    if site not in _seen:
        _seen.add(site)
        warnings.warn(msg, RuntimeWarning, stacklevel=3)
