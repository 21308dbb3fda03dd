"""Integer codes of the task-program and model-blob encodings, parsed from ``include/b200sim_codes.h`` and
``include/b200sim_model.h`` (the single source of truth)."""

from __future__ import annotations

import re
from pathlib import Path

_INCLUDE = Path(__file__).resolve().parent.parent / "include"
_HEADERS = (_INCLUDE / "b200sim_codes.h", _INCLUDE / "b200sim_model.h")


def _parse() -> dict[str, int]:
    out: dict[str, int] = {}
    for header in _HEADERS:
        for line in header.read_text().splitlines():
            m = re.match(r"#define\s+([A-Z0-9_]+)\s+(-?(?:0x[0-9A-Fa-f]+|\d+))\b", line)
            if m:
                out[m.group(1)] = int(m.group(2), 0)
    return out


CODES = _parse()
globals().update(CODES)


def __getattr__(name: str) -> int:  # pragma: no cover - only hit on typos
    raise AttributeError(f"unknown code {name}")
