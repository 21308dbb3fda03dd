"""Compiles the reference's robot MJCFs with ksim_b200.mjcf and stores the COMPILED arrays as assets.

The GPU box has no /root/reference, so the benchmark models travel as compiled ``.npz`` fixtures (arrays named
after mjModel fields + names + options); this script is how they were made:

    python scripts/compile_reference_models.py            # needs /root/reference

Sources: examples/data/scene.mjcf (humanoid, BASELINE configs 1/2/4);
examples/kbot/robot/kbot-headless/robot.mjcf + metadata.json (K-Bot, configs 3/5).  The K-Bot scene comes from the
un-vendored ``mujoco_scenes`` package in the reference; here the robot is placed on a plain floor plane
(``contype=1``, default friction), which is what its "smooth" scene amounts to for the contact model.
"""

from __future__ import annotations

import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))

import json  # noqa: E402
import tempfile  # noqa: E402
import xml.etree.ElementTree as ET  # noqa: E402

from ksim_b200.mjcf import load_mjcf, save_model  # noqa: E402

REFERENCE = Path("/root/reference")


def main() -> None:
    out = ROOT / "ksim_b200" / "assets"
    out.mkdir(exist_ok=True)
    humanoid = load_mjcf(REFERENCE / "examples" / "data" / "scene.mjcf")
    save_model(humanoid, out / "humanoid.npz")
    print("humanoid:", humanoid.nq, humanoid.nv, humanoid.nbody, "->", out / "humanoid.npz")
    kdir = REFERENCE / "examples" / "kbot" / "robot" / "kbot-headless"
    tree = ET.parse(str(kdir / "robot.mjcf"))
    world = tree.getroot().find("worldbody")
    floor = ET.Element("geom", dict(name="floor", type="plane", size="0 0 0.05", contype="1", conaffinity="0"))
    world.insert(0, floor)
    with tempfile.TemporaryDirectory() as tmp:
        path = Path(tmp) / "kbot_scene.mjcf"
        tree.write(str(path))
        kbot = load_mjcf(path)
    save_model(kbot, out / "kbot.npz")
    meta = json.loads((kdir / "metadata.json").read_text())
    (out / "kbot_metadata.json").write_text(json.dumps(meta, indent=1, sort_keys=True))
    print("kbot:", kbot.nq, kbot.nv, kbot.nbody, kbot.nu, "pairs", len(kbot.arrays["pair_geom1"]), "->", out / "kbot.npz")


if __name__ == "__main__":
    main()
