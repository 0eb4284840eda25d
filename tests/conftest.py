import importlib.util
import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def load_package():
    """Import go-mp3_b200/ (hyphenated directory) as module go_mp3_b200."""
    if "go_mp3_b200" in sys.modules:
        return sys.modules["go_mp3_b200"]
    spec = importlib.util.spec_from_file_location("go_mp3_b200", os.path.join(ROOT, "go-mp3_b200", "__init__.py"))
    mod = importlib.util.module_from_spec(spec)
    sys.modules["go_mp3_b200"] = mod
    spec.loader.exec_module(mod)
    return mod


@pytest.fixture(scope="session")
def pkg():
    return load_package()


@pytest.fixture(scope="session")
def fixtures_dir():
    return os.path.join(ROOT, "tests", "golden", "fixtures")


@pytest.fixture(scope="session")
def classic_lame(fixtures_dir):
    with open(os.path.join(fixtures_dir, "classic_lame.mp3"), "rb") as f:
        return f.read()


@pytest.fixture(scope="session")
def mpeg2(fixtures_dir):
    with open(os.path.join(fixtures_dir, "mpeg2.mp3"), "rb") as f:
        return f.read()
