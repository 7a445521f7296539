"""`-m gpu`: the rttest-style suite (tests/rttest_cuda.py) — every builtin scene of the reference's tests/tests.toml plus
three glTF fixtures rendered through the `--backend cuda` driver into EXR files and compared with CPU-blessed
references: rttest's MSE / max-diff metric, FLIP against the CPU-vs-CPU noise floor, mean luminance within 3 sigma."""
import pytest

pytestmark = pytest.mark.gpu


def test_rttest_suite_against_cpu_blessed_references(rc, oracle, tmp_path):
    import rttest_cuda
    results = rttest_cuda.run_suite(rc, oracle, ("-s", "4", "-l", "1"), workdir=str(tmp_path))
    assert len(results) == len(rttest_cuda.SUITE) + len(rttest_cuda.GLTF_SUITE)
    failed = [r for r in results if not r["passed"]]
    assert not failed, failed
    for r in results:
        if "luminance_z_vs_other_seed" in r:
            assert r["luminance_z_vs_other_seed"] <= 3.0 or r["mse_beauty"] <= r["mse_noise_floor"], r
