"""GPU parity for HaarSeg (csrc/haar.cu through the C ABI): breakpoints bit-exact, means 1e-9,
against the reference's vectors (tests/golden/haar.npz) and the C oracle on seeded inputs."""
import numpy as np
import pytest

from oracle import haar as OH

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def H():
    from cnvkit_b200.segmentation import haar

    return haar


def per_arm_starts(arm, lo, off):
    return (lo - off[arm]).tolist()


def test_reference_fixture_per_arm(H, golden):
    g = golden("haar")
    x, w, off = g["cnr__log2"], g["cnr__weight"], g["arm_offsets"]
    for key, wts, q in (("w", w, 1e-4), ("u", None, 1e-4), ("q05", w, 0.05)):
        arm, lo, hi, mean = H.segment_arms(x, wts, off, q)
        np.testing.assert_array_equal(per_arm_starts(arm, lo, off), g[f"{key}_seg_start"])
        np.testing.assert_array_equal(np.bincount(arm, minlength=len(off) - 1), np.diff(g[f"{key}_seg_offsets"]))
        if key != "q05":
            np.testing.assert_allclose(mean, g[f"{key}_seg_mean"], rtol=1e-9, atol=1e-12)


def test_planted_steps_kat(H, golden):
    # test/test_haar.py:219-230: three segments with means ~ [0, -2, -2.3]
    g = golden("haar")
    r = H.haarSeg(g["kat_signal"], 1e-4)
    np.testing.assert_array_equal(r["start"], g["kat_start"])
    np.testing.assert_allclose(r["mean"], g["kat_mean"], rtol=1e-9)
    assert len(r["start"]) == 3


def test_constant_and_tiny_arms(H):
    # test/test_haar.py:200-207, 232-235: constant signal -> one segment; tiny arms -> one segment
    x = np.concatenate([np.full(200, 1.25), [0.3], [0.1, 0.2], np.zeros(3), np.arange(40.0) % 2])
    off = np.array([0, 200, 201, 203, 206, 246])
    arm, lo, hi, mean = H.segment_arms(x, None, off, 1e-4)
    o = OH.segment_arms(x, None, off, 1e-4)
    np.testing.assert_array_equal(arm, o[0])
    np.testing.assert_array_equal(lo, o[1])
    np.testing.assert_array_equal(hi, o[2])
    assert (arm == 0).sum() == 1 and mean[0] == 1.25


@pytest.mark.parametrize("weighted", [False, True])
@pytest.mark.parametrize("q", [1e-4, 1e-2, 0.3])
def test_vs_oracle_random_arms(H, weighted, q):
    rng = np.random.default_rng(17 + int(q * 1000) + weighted)
    sizes = [1, 2, 3, 5, 31, 32, 33, 64, 65, 100, 400, 1023, 1025, 5000, 20000]
    xs = []
    for n in sizes:
        x = rng.normal(0, 0.3, n)
        for _ in range(max(1, n // 300)):
            a = int(rng.integers(0, max(1, n - 2)))
            b = min(n, a + int(rng.integers(2, 200)))
            x[a:b] += rng.choice([-1.0, 0.6, 1.0])
        xs.append(np.round(x, 3) if n % 2 else x)  # ties / plateaus on odd sizes
    x = np.concatenate(xs)
    w = rng.uniform(0.1, 1.0, len(x)) if weighted else None
    off = np.concatenate([[0], np.cumsum(sizes)])
    arm, lo, hi, mean = H.segment_arms(x, w, off, q)
    o_arm, o_lo, o_hi, o_mean = OH.segment_arms(x, w, off, q)
    np.testing.assert_array_equal(arm, o_arm)
    np.testing.assert_array_equal(lo, o_lo)
    np.testing.assert_array_equal(hi, o_hi)
    np.testing.assert_allclose(mean, o_mean, rtol=1e-9, atol=1e-12)


def test_nan_weights_segment_means(H):
    # test/test_haar.py:345-362: NaN weights are excluded from the weighted mean
    x = np.concatenate([np.zeros(50), np.ones(50)])
    w = np.ones(100)
    w[10] = np.nan
    arm, lo, hi, mean = H.segment_arms(x, None, np.array([0, 100]), 1e-4)
    assert list(lo) == [0, 50]
    from oracle import haar as OH2
    _, m = OH2.segment(x, None, 1e-4)
    np.testing.assert_allclose(mean, m)


def test_wgs_scale_properties(H):
    """C3-sized: 3M bins / 48 arms; planted steps are found, the table tiles every arm, and a
    second call is identical (deterministic)."""
    from cnvkit_b200 import synth

    d = synth.wgs_sample(total_bins=3_000_000, quantise=False)
    arm, lo, hi, mean = H.segment_arms(d["log2"], d["weight"], d["arm_offsets"], 1e-4)
    off = d["arm_offsets"]
    for a in range(len(off) - 1):
        sel = arm == a
        assert lo[sel][0] == off[a] and hi[sel][-1] == off[a + 1] and (lo[sel][1:] == hi[sel][:-1]).all()
    found = set(hi.tolist())
    hits = sum(1 for (a, b, step) in d["truth"] if abs(step) >= 0.58 and b - a >= 16
               and any(abs(a - f) <= 2 for f in found if abs(a - f) <= 2))
    big = sum(1 for (a, b, step) in d["truth"] if abs(step) >= 0.58 and b - a >= 16)
    assert hits >= 0.9 * big
    again = H.segment_arms(d["log2"], d["weight"], d["arm_offsets"], 1e-4)
    np.testing.assert_array_equal(lo, again[1])
    np.testing.assert_array_equal(mean, again[3])
