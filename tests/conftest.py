import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def _run(cmd, **kw):
    r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, **kw)
    if r.returncode != 0:
        raise RuntimeError("command failed: %s\n%s" % (" ".join(cmd), r.stdout))
    return r.stdout


@pytest.fixture(scope="session")
def built():
    """Builds the host library, the oracle port and (when /root/reference exists) oracle/_ref."""
    _run(["make", "-s", "-C", os.path.join(ROOT, "methyldackel_b200", "csrc"), "host"])
    _run(["make", "-s", "-C", os.path.join(ROOT, "oracle"), "port"])
    _run(["make", "-s", "-C", os.path.join(ROOT, "oracle"), "ref"])
    _run(["make", "-s", "-C", os.path.join(ROOT, "tests", "native")])
    return {
        "ref_bin": os.path.join(ROOT, "oracle", "_ref", "MethylDackel"),
        "mdsynth": os.path.join(ROOT, "methyldackel_b200", "lib", "mdsynth"),
        "fixtures": os.path.join(ROOT, "tests", "golden", "fixtures"),
    }


@pytest.fixture(scope="session")
def synth(built, tmp_path_factory):
    """Factory: synth(name, *mdsynth args) -> prefix of a cached synthetic data set."""
    base = tmp_path_factory.mktemp("synth")
    cache = {}

    def make(name, *args):
        if name not in cache:
            prefix = str(base / name)
            _run([built["mdsynth"], "--out", prefix] + [str(a) for a in args])
            cache[name] = prefix
        return cache[name]
    return make
