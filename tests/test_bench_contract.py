"""bench.py's reference arm and JSON contract, on the smallest workload (CPU only)."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_prints_one_json_line():
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--workload", "small",
                          "--steps", "1", "--warmup", "0"], capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [ln for ln in out.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["unit"] == "GVoxels/s" and d["higher_is_better"] is True
    assert d["value"] > 0 and d["cpu_baseline"]["kind"] == "reference" and d["cpu_baseline"]["cores"] >= 1
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert d["metric"] == "gaussian_filter_gvoxels_per_s" and "workload" in d["config"] and d["gpu_launches"] == 0


def test_reference_arm_scipy_per_chunk_form():
    """B200_REFERENCE_ARM=scipy (and any box without a copy of the reference): the same scipy call per halo'd chunk
    with the halo depth restated by the oracle."""
    env = dict(os.environ, B200_REFERENCE_ARM="scipy")
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--workload", "small",
                          "--steps", "1", "--warmup", "0"], capture_output=True, text=True, timeout=600, cwd=ROOT,
                         env=env)
    assert out.returncode == 0, out.stderr[-2000:]
    d = json.loads(out.stdout.strip().splitlines()[-1])
    assert d["value"] > 0 and "dask_image.ndfilters" not in d["config"]["sample"]


def test_reference_arm_other_ranks_exit_quietly():
    env = dict(os.environ, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1")
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2",
                          "--workload", "small", "--steps", "1", "--warmup", "0"], capture_output=True, text=True,
                         timeout=600, cwd=ROOT, env=env)
    assert out.returncode == 0, out.stderr[-2000:]
    assert out.stdout.strip() == ""


def test_reference_arm_runs_the_reference_api_when_present():
    """With a copy of the reference at hand (baseline/_ref made by build(), or /root/reference) the CPU arm calls the
    unmodified dask_image.ndfilters.gaussian_filter over a thread-pool map_overlap; the result is the scipy-per-chunk
    form's, bit for bit, and whole-array scipy's."""
    import numpy as np
    import pytest
    import scipy.ndimage as ndi

    code = (
        "import sys, numpy as np\n"
        "sys.path.insert(0, %r)\n"
        "from oracle import reference_arm, oracle as orc\n"
        "import scipy.ndimage as ndi\n"
        "ref = reference_arm.reference_ndfilters()\n"
        "if ref is None:\n"
        "    print('ABSENT'); raise SystemExit(0)\n"
        "a = np.random.default_rng(1).random((40, 44, 36), dtype=np.float32)\n"
        "got = reference_arm.gaussian_filter(a, (16, 16, 16), 4, sigma=(2.0, 2.0, 2.0), mode='reflect')\n"
        "alt = orc.map_overlap_emulated(ndi.gaussian_filter, a, (16,) * 3, orc.gaussian_depth(2.0, 4.0, 3), 'none',\n"
        "                               n_threads=4, sigma=2.0, mode='reflect')\n"
        "assert got.dtype == a.dtype and np.array_equal(got, alt) and np.array_equal(got, ndi.gaussian_filter(a, 2.0))\n"
        "assert 'dask_image_b200' not in sys.modules and not any('b200' in m for m in sys.modules)\n"
        "print('OK', ref.__file__)\n" % ROOT)
    out = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    if "ABSENT" in out.stdout:
        pytest.skip("no copy of the reference present")
    assert "OK" in out.stdout
    del np, ndi
