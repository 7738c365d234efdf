"""Bridge to the host application's ``.dat`` parser (main.py:59-180).

The tunnel's input producer is the reference backend's own
``parse_dat_file(path) -> (coords, fixes)``; this package deliberately does not
re-implement its repair logic (Lednicer merge, winding fix, LE de-duplication).
When the tunnel is deployed beside ``main.py`` that function is imported from
there.  ``main.py`` imports ``slowapi`` at module level (main.py:11-13); if that
package is absent a no-op stand-in is registered first so the two pure parser
functions can still be used.

``read_plain_dat`` is a minimal two-column reader for stand-alone use (already
clean Selig files); it performs no repairs.
"""
from __future__ import annotations

import importlib
import os
import sys
import types
from typing import Callable, Optional


def _stub_slowapi() -> None:
    try:
        importlib.import_module("slowapi")
        return
    except ImportError:
        pass
    pkg = types.ModuleType("slowapi")
    util = types.ModuleType("slowapi.util")
    errors = types.ModuleType("slowapi.errors")

    class Limiter:
        def __init__(self, *a, **k):
            pass

        def limit(self, *a, **k):
            return lambda fn: fn

    class RateLimitExceeded(Exception):
        pass

    pkg.Limiter = Limiter
    pkg._rate_limit_exceeded_handler = lambda *a, **k: None
    util.get_remote_address = lambda request: "0.0.0.0"
    errors.RateLimitExceeded = RateLimitExceeded
    sys.modules.setdefault("slowapi", pkg)
    sys.modules.setdefault("slowapi.util", util)
    sys.modules.setdefault("slowapi.errors", errors)


def load_reference_parser(app_dir: Optional[str] = None) -> Callable:
    """Import ``parse_dat_file`` from the application's ``main.py``.

    ``app_dir`` defaults to ``$AEROLAB_APP_DIR``; if neither is given ``main``
    must already be importable.
    """
    app_dir = app_dir or os.environ.get("AEROLAB_APP_DIR")
    if app_dir and app_dir not in sys.path:
        sys.path.insert(0, app_dir)
    _stub_slowapi()
    mod = importlib.import_module("main")
    return mod.parse_dat_file


def read_plain_dat(path: str):
    """Two numeric columns per line; non-numeric lines are skipped.  No repairs."""
    coords = []
    skipped = 0
    with open(path, "r") as fh:
        for line in fh:
            parts = line.replace(",", " ").split()
            try:
                if len(parts) >= 2:
                    coords.append([float(parts[0]), float(parts[1])])
                    continue
            except ValueError:
                pass
            if line.strip():
                skipped += 1
    if len(coords) < 10:
        raise ValueError(f"{path}: fewer than 10 coordinate pairs")
    fixes = [f"Non-coordinate lines skipped: {skipped}"] if skipped else []
    return coords, fixes


def resolve_parser(parser: Optional[Callable]) -> Callable:
    if parser is not None:
        return parser
    try:
        return load_reference_parser()
    except ImportError as e:
        raise ImportError(
            "load_dat() needs the host application's parse_dat_file (main.py:59); put main.py on "
            "sys.path, set AEROLAB_APP_DIR, or pass parser=... (e.g. aerolab_lbm.dat.read_plain_dat)"
        ) from e
