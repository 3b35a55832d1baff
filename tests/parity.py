"""Frame comparison used by the GPU parity tests.

The bar: every pixel within 1 LSB per 8-bit channel (8-bit = the reference's truncating
Canvas::scale_color, canvas.rs:39-43) on >= 99.9 % of pixels, a bounded number of outliers elsewhere (rays whose
threshold tests — silhouettes, pattern edges, shadow boundaries — flip on a last-place difference), and the
never-rendered last row / column black (camera.rs:80-81)."""
import numpy as np

MIN_FRACTION_WITHIN_1LSB = 0.999
MAX_FRACTION_GROSS = 0.0005  # pixels off by more than 8 LSB (a flipped threshold), bounded separately


def compare_frames(got_u8, want_u8, got_f32=None, want_f32=None):
    assert got_u8.shape == want_u8.shape
    d = np.abs(got_u8.astype(np.int16) - want_u8.astype(np.int16)).max(axis=2)
    n = d.size
    rep = {
        "pixels": int(n),
        "exact_u8": float((d == 0).sum() / n),
        "within_1lsb": float((d <= 1).sum() / n),
        "gross": float((d > 8).sum() / n),
        "max_u8_diff": int(d.max()),
    }
    if got_f32 is not None and want_f32 is not None:
        rep["bit_exact_f32"] = float((got_f32.view(np.uint32) == want_f32.view(np.uint32)).all(axis=2).sum() / n)
        with np.errstate(invalid="ignore"):
            rep["max_abs_f32"] = float(np.nanmax(np.abs(got_f32 - want_f32)))
    return rep


def assert_parity(rep, min_within=MIN_FRACTION_WITHIN_1LSB, max_gross=MAX_FRACTION_GROSS, label=""):
    assert rep["within_1lsb"] >= min_within, f"{label}: only {rep['within_1lsb']:.5f} of pixels within 1 LSB: {rep}"
    assert rep["gross"] <= max_gross, f"{label}: {rep['gross']:.5f} of pixels off by more than 8 LSB: {rep}"


def assert_unrendered_border(u8, f32=None):
    assert not u8[-1, :, :].any() and not u8[:, -1, :].any(), "last row / column must stay black (camera.rs:80-81)"
    if f32 is not None:
        assert not f32[-1, :, :].any() and not f32[:, -1, :].any()
