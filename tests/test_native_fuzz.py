"""CPU: sanitizer fuzzers of the GPU-free host code (tests/native/*.cpp): compiled with g++ -fsanitize=address,undefined
against csrc/factors.cpp resp. csrc/plan.cpp and run; any out-of-bounds access, overflow or wrong result fails."""
import os
import shutil
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CSRC = os.path.join(ROOT, "eigen-git-mirror_b200", "csrc")


@pytest.mark.timeout(600)
@pytest.mark.parametrize("name,unit", [("fuzz_factors", "factors.cpp"), ("fuzz_plan", "plan.cpp")])
def test_host_code_under_sanitizers(name, unit, tmp_path):
    gxx = "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else shutil.which("g++")
    if not gxx:
        pytest.skip("no g++")
    exe = str(tmp_path / name)
    cmd = [gxx, "-O1", "-g", "-std=c++17", "-pthread", "-fsanitize=address,undefined", "-fno-omit-frame-pointer",
           "-fno-sanitize-recover=undefined", "-I", CSRC, os.path.join(ROOT, "tests", "native", name + ".cpp"),
           os.path.join(CSRC, unit), "-o", exe]
    env = dict(os.environ)
    env.pop("CXX", None)
    build = subprocess.run(cmd, capture_output=True, text=True, env=env)
    if build.returncode != 0 and "asan" in (build.stderr + build.stdout).lower():
        pytest.skip("sanitizer runtime not installed")
    assert build.returncode == 0, build.stderr[-3000:]
    run = subprocess.run([exe], capture_output=True, text=True, timeout=500)
    assert run.returncode == 0 and "done, bad=0" in run.stdout, run.stdout[-2000:] + run.stderr[-4000:]
