"""f-2: the reference's susceptibility table (corner_contribution/magnet.cpp) from fused magnetisations."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLD = os.path.join(ROOT, "tests", "golden")
TOOL = os.path.join(ROOT, "blume_capel_b200", "host", "bc_magnet")


def test_table_is_byte_identical_to_the_reference_tool(tmp_path):
    """magnet_expected.csv was written by the unmodified magnet.cpp from cluster files of the same samples
    (tests/golden/make_magnet_golden.py); bc_magnet must print the same bytes from the magnetisations."""
    if not os.access(TOOL, os.X_OK):
        pytest.skip("bc_magnet not built (needs libbc_engine.so)")
    out = tmp_path / "m.csv"
    subprocess.run([TOOL, str(out), "--from-m", os.path.join(GOLD, "magnet_samples.txt")], check=True)
    assert out.read_bytes() == open(os.path.join(GOLD, "magnet_expected.csv"), "rb").read()


@pytest.mark.gpu
def test_bc_magnet_on_gpu(tmp_path):
    out = tmp_path / "chi.csv"
    subprocess.run([TOOL, str(out), "0.608", "1.966", "1.0", "--L", "32,64", "--runs", "20", "--samples", "30",
                    "--burn-sweeps", "2000", "--sweeps-per-sample", "200", "--seed", "5"], check=True)
    rows = out.read_text().strip().splitlines()
    assert rows[0] == "batch,L,magnetic_susceptibility,standard_error"
    assert len(rows) == 1 + 2 * 2
    for r in rows[1:]:
        b, L, chi, err = r.split(",")
        assert int(L) in (32, 64) and float(chi) > 0 and float(err) >= 0
