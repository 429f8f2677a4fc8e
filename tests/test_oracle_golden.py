"""The oracle's own restatement of the shuffle argument (oracle/shuffle_ref.py) against the golden
vectors produced by the UNMODIFIED reference (tests/golden/, oracle/gen_golden.py): identical
proof bytes under the same seed, identical verdicts on honest and corrupted inputs."""
import os

import pytest

import shuffle_cases as sc
from oracle import ark_surface


@pytest.fixture()
def c_surface():
    ark_surface.set_backend("c")
    yield ark_surface
    ark_surface.set_backend("py")


@pytest.mark.parametrize("name", ["shuffle_N8_seed1234.json", "shuffle_N16_seed77.json", "shuffle_N64_seed2024.json", "shuffle_N128_seed4096.json"])
def test_prove_bytes(c_surface, name):
    sc.check_prove_matches_golden(c_surface, sc.load_case(name))


@pytest.mark.parametrize("name", ["shuffle_N8_seed1234.json", "shuffle_N64_seed2024.json"])
def test_verify_verdicts(c_surface, name):
    sc.check_verify_matches_golden(c_surface, sc.load_case(name))


def test_python_backend_small():
    ark_surface.set_backend("py")
    sc.check_prove_matches_golden(ark_surface, sc.load_case("shuffle_N8_seed1234.json"))


@pytest.mark.skipif(not os.path.isdir("/root/reference/curdleproofs"), reason="reference not mounted (GPU box)")
def test_fixtures_are_reproducible_from_the_reference(tmp_path):
    """Re-run the generator's N=8 case against the mounted reference and compare with the fixture."""
    import json
    import subprocess
    import sys

    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    code = (
        "import sys, json; sys.argv=['x']; sys.path.insert(0, %r)\n"
        "import importlib.util\n"
        "spec = importlib.util.spec_from_file_location('gg', %r); m = importlib.util.module_from_spec(spec); spec.loader.exec_module(m)\n"
        "print(json.dumps(m.one_case(1234, 8)['proof']))\n"
    ) % (root, os.path.join(root, "oracle", "gen_golden.py"))
    out = subprocess.check_output([sys.executable, "-c", code], cwd=root).decode().strip().splitlines()[-1]
    assert json.loads(out) == sc.load_case("shuffle_N8_seed1234.json")["proof"]
