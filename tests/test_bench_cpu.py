"""bench.py contract, the part that runs without a GPU: the reference arm (`--impl reference`) times
the unmodified reference (oracle/_ref/iSS.e) on the host cores and prints the JSON line the driver
parses; the engine arm must refuse to run without the CUDA library rather than fall back."""
import json
import os
import subprocess
import sys

import pytest

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.skipif(not os.path.exists(os.path.join(REPO, "oracle", "_ref", "iSS.e")),
                    reason="oracle/_ref not built (python __graft_entry__.py builds it next to /root/reference)")
def test_reference_arm_prints_the_contract_line():
    r = subprocess.run([sys.executable, os.path.join(REPO, "bench.py"), "--impl", "reference", "--steps", "1",
                        "--warmup", "0"], capture_output=True, text=True, timeout=600, cwd=REPO)
    assert r.returncode == 0, r.stderr[-2000:]
    line = json.loads(r.stdout.strip().splitlines()[-1])
    assert line["impl"] == "reference"
    assert line["metric"] == "sampled_hadrons_per_sec" and line["unit"] == "hadrons/s"
    assert line["higher_is_better"] is True and line["n_gpus"] == 1 and line["steps"] == 1
    assert line["value"] > 1e4
    assert line["config"]["workload_key"] == "c4" and "C4" in line["config"]["workload"]
    cb = line["cpu_baseline"]
    assert cb["kind"] == "reference" and cb["cores"] >= 1 and cb["value"] == line["value"]
    assert "oracle/_ref/iSS.e" in cb["sample"]
    e2e = line["e2e"]
    assert e2e["value"] == line["value"] and e2e["unit"] == line["unit"]
    assert e2e["h2d_bytes_per_step"] == 0 and e2e["d2h_bytes_per_step"] == 0


def test_reference_arm_on_other_ranks_exits_quietly():
    """under torchrun only rank 0 runs the reference arm; the others exit 0 without output"""
    env = dict(os.environ, RANK="1", LOCAL_RANK="1", WORLD_SIZE="2")
    r = subprocess.run([sys.executable, os.path.join(REPO, "bench.py"), "--impl", "reference", "--gpus", "2",
                        "--steps", "1", "--warmup", "0"], capture_output=True, text=True, timeout=120,
                       cwd=REPO, env=env)
    assert r.returncode == 0
    assert r.stdout.strip() == ""


def _no_gpu():
    try:
        import torch
        return not torch.cuda.is_available()
    except Exception:
        return True


NO_GPU_FACADE = """
import sys, tempfile
sys.path.insert(0, %(repo)r); sys.path.insert(0, %(tests)r)
import cases
from iss_b200 import capi
g = cases.load(cases.ONE_CELL[0])
d = tempfile.mkdtemp()
param, surf, over = cases.materialise(g, d)
over.update(number_of_repeated_sampling=5, perform_checks=0, use_OSCAR_format=0)
s = capi.Sampler(d, param, surf, table_path=cases.tables_for(g), **over)
assert s.read_in_FO_surface() == 0
s.set_random_seed(1)
print("generate_samples returned", s.generate_samples())
"""


@pytest.mark.skipif(not _no_gpu(), reason="needs a host without a GPU")
def test_no_cpu_fallback_without_a_gpu(tmp_path):
    """the product path has no CPU fallback: the C ABI refuses to make a handle and class iSS ends the
    run the way the reference ends on a fatal error (message + exit(-1)), it does not sample on the host"""
    import ctypes as C
    sys.path.insert(0, REPO)
    from iss_b200 import capi
    L = capi.cuda_lib()
    h = C.c_void_p()
    assert L.iss_cuda_create(0, C.byref(h)) == 1        # ISS_ERR_CUDA
    assert not h.value
    script = tmp_path/"facade.py"
    script.write_text(NO_GPU_FACADE % dict(repo=REPO, tests=os.path.join(REPO, "tests")))
    r = subprocess.run([sys.executable, str(script)], capture_output=True, text=True, timeout=300)
    assert r.returncode != 0
    assert "generate_samples returned" not in r.stdout
    assert "no usable CUDA device" in (r.stdout + r.stderr)
