"""Shared fixtures.  GPU tests are marked ``gpu`` (run with ``-m gpu`` on a B200); everything else runs on CPU."""

from __future__ import annotations

import sys
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent
if str(ROOT) not in sys.path:
    sys.path.insert(0, str(ROOT))


def pytest_configure(config: pytest.Config) -> None:
    config.addinivalue_line("markers", "gpu: needs a CUDA device (B200); run with -m gpu")


@pytest.fixture(scope="session")
def walking_spec():  # noqa: ANN201
    from ksim_b200.examples.walking import make_spec

    return make_spec()


@pytest.fixture(scope="session")
def walking_program(walking_spec):  # noqa: ANN001, ANN201
    from ksim_b200.program import build_program

    return build_program(walking_spec)


@pytest.fixture(scope="session")
def walking_randomizers(walking_spec):  # noqa: ANN001, ANN201
    from ksim_b200.examples.walking import RANDOMIZERS

    return {k: f(walking_spec.model) for k, f in RANDOMIZERS.items()}


@pytest.fixture(scope="session")
def kbot_spec():  # noqa: ANN201
    from ksim_b200.examples.kbot import make_spec

    return make_spec()


@pytest.fixture(scope="session")
def kbot_randomizers(kbot_spec):  # noqa: ANN001, ANN201
    from ksim_b200.examples.kbot import RANDOMIZERS

    return {k: f(kbot_spec.model) for k, f in RANDOMIZERS.items()}


@pytest.fixture(scope="session")
def oracle_built() -> None:
    import oracle

    oracle.build()
