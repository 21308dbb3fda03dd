"""Integer codes of the task-program encoding, parsed from ``include/b200sim_codes.h`` (single source of truth)."""

from __future__ import annotations

import re
from pathlib import Path

_HEADER = Path(__file__).resolve().parent.parent / "include" / "b200sim_codes.h"


def _parse() -> dict[str, int]:
    out: dict[str, int] = {}
    for line in _HEADER.read_text().splitlines():
        m = re.match(r"#define\s+([A-Z0-9_]+)\s+(-?(?:0x[0-9A-Fa-f]+|\d+))\b", line)
        if m:
            out[m.group(1)] = int(m.group(2), 0)
    return out


CODES = _parse()
globals().update(CODES)


def __getattr__(name: str) -> int:  # pragma: no cover - only hit on typos
    raise AttributeError(f"unknown code {name}")
