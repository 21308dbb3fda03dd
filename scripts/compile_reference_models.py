"""Compiles the reference's robot MJCFs with ksim_b200.mjcf and stores the COMPILED arrays as assets.

The GPU box has no /root/reference, so the benchmark models travel as compiled ``.npz`` fixtures (arrays named
after mjModel fields + names + options); this script is how they were made:

    python scripts/compile_reference_models.py            # needs /root/reference

Sources: examples/data/scene.mjcf (humanoid, BASELINE configs 1/2/4).
"""

from __future__ import annotations

import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))

from ksim_b200.mjcf import load_mjcf, save_model  # noqa: E402

REFERENCE = Path("/root/reference")


def main() -> None:
    out = ROOT / "ksim_b200" / "assets"
    out.mkdir(exist_ok=True)
    humanoid = load_mjcf(REFERENCE / "examples" / "data" / "scene.mjcf")
    save_model(humanoid, out / "humanoid.npz")
    print("humanoid:", humanoid.nq, humanoid.nv, humanoid.nbody, "->", out / "humanoid.npz")


if __name__ == "__main__":
    main()
