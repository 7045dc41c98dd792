"""Host-side file-format code under AddressSanitizer + UBSan: formats.hpp / reorder.hpp parse mutated granne files
(truncated, bit-flipped, corrupted headers / offset chunks) without any out-of-bounds access.  CPU only."""
import os
import shutil
import subprocess

import pytest

from helpers.data import build_fixture

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def harness(tmp_path_factory):
    if shutil.which("g++") is None:
        pytest.skip("g++ not available")
    exe = str(tmp_path_factory.mktemp("fuzz") / "formats_fuzz")
    subprocess.check_call(["g++", "-std=c++17", "-O1", "-g", "-fsanitize=address,undefined", "-fno-sanitize-recover=all",
                           "-I" + os.path.join(ROOT, "granne_b200", "csrc"),
                           os.path.join(ROOT, "tests", "helpers", "formats_fuzz.cpp"), "-o", exe])
    return exe


@pytest.mark.parametrize("kind,code,n,dim", [("angular", 0, 300, 8), ("angular_int", 1, 130, 16), ("embeddings", 2, 200, 6)])
def test_mutated_files_never_read_out_of_bounds(harness, oracle, tmp_path, kind, code, n, dim):
    el, g, ib, eb, mb = build_fixture(oracle, kind, n, dim, seed=3, num_neighbors=8, max_search=10, layer_multiplier=5.0)
    paths = {}
    for name, data in (("index", ib), ("elements", eb), ("embeddings", mb)):
        if data is not None:
            paths[name] = str(tmp_path / name)
            with open(paths[name], "wb") as f:
                f.write(data)
    r = subprocess.run([harness, paths["index"], paths["elements"], str(code), paths.get("embeddings", "-"), "4000", "11"],
                       capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr[-3000:]
    parsed, rejected = (int(x) for x in r.stdout.split()[1::2])
    assert parsed > 500 and rejected > 500  # both the accept and the reject paths were exercised
