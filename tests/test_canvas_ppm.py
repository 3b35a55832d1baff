"""canvas.rs:58-200 (to_ppm, canvas_from_ppm) and uv.rs:346-377 (UVImage): the reference's own test vectors
(canvas.rs:227-399, uv.rs:640-672), run against BOTH restatements — the CPU oracle and the product's host library
(native C++ reader / writer; no GPU needed) — plus byte equality of the two on random canvases."""
import numpy as np
import pytest

from tests.helpers import assert_abs_diff_eq, assert_eq


@pytest.fixture(scope="module")
def host():
    import ray_tracer_challenge_b200 as rt

    return rt.new_session()


@pytest.fixture(params=["oracle", "host"])
def lib(request, oracle, host):
    return oracle if request.param == "oracle" else host


def test_ppm_header(lib):  # canvas.rs:243-251
    lines = lib.new_canvas(20, 5).to_ppm().splitlines()
    assert lines[:3] == ["P3", "20 5", "255"]


def test_ppm_pixel_data(lib):  # canvas.rs:253-271 — 0.5 -> 127: truncation, not rounding
    c = lib.new_canvas(5, 3)
    c.write_pixel(0, 0, (1.5, 0, 0))
    c.write_pixel(2, 1, (0, 0.5, 0))
    c.write_pixel(4, 2, (-0.5, 0, 1))
    lines = c.to_ppm().splitlines()[3:]
    assert lines == ["255 0 0 0 0 0 0 0 0 0 0 0 0 0 0", "0 0 0 0 0 0 0 127 0 0 0 0 0 0 0", "0 0 0 0 0 0 0 0 0 0 0 0 0 0 255"]


def test_splitting_long_ppm_lines(lib):  # canvas.rs:273-307
    c = lib.new_canvas(10, 2)
    c.data[:] = (1, 0.8, 0.6)
    ppm = c.to_ppm()
    lines = ppm.splitlines()[3:]
    assert lines == ["255 204 153 255 204 153 255 204 153 255 204 153 255 204 153 255 204",
                     "153 255 204 153 255 204 153 255 204 153 255 204 153"] * 2
    assert ppm.endswith("\n") and max(len(l) for l in ppm.splitlines()) <= 70


def test_reading_file_with_wrong_magic_number(lib):  # canvas.rs:309-322
    from ray_tracer_challenge_b200.api import RtcError

    with pytest.raises(RtcError, match="IncorrectFormat.*Incorrect magic number"):
        lib.canvas_from_ppm("P32\n        1 1\n        255\n        0 0 0")


def test_malformed_headers(lib):  # ParseError::MalformedDimensionHeader / ParseIntError (canvas.rs:98-117,139-153)
    from ray_tracer_challenge_b200.api import RtcError

    with pytest.raises(RtcError, match="MalformedDimensionHeader"):
        lib.canvas_from_ppm("P3\n1 1 1\n255\n0 0 0\n")
    with pytest.raises(RtcError, match="ParseIntError"):
        lib.canvas_from_ppm("P3\n1 x\n255\n0 0 0\n")
    with pytest.raises(RtcError, match="ParseIntError"):
        lib.canvas_from_ppm("P3\n1 1\n255 0\n0 0 0\n")  # the whole third line is the scale
    with pytest.raises(RtcError, match="ParseIntError"):
        lib.canvas_from_ppm("P3\n1 1\n255\n0 -1 0\n")


def test_reading_ppm_returns_canvas_with_correct_size(lib):  # canvas.rs:324-337
    rows = "0 0 0  0 0 0  0 0 0  0 0 0  0 0 0\n" * 4
    c = lib.canvas_from_ppm("P3\n        10 2\n        255\n" + rows)
    assert (c.width, c.height) == (10, 2)


def test_reading_pixel_data_from_ppm_file(lib):  # canvas.rs:339-367
    c = lib.canvas_from_ppm("""P3
        4 3
        255
        255 127 0  0 127 255  127 255 0  255 255 255
        0 0 0  255 0 0  0 255 0  0 0 255
        255 255 0  0 255 255  255 0 255  127 127 127""")
    h = 0.49803922
    want = {(0, 0): (1, h, 0), (1, 0): (0, h, 1), (2, 0): (h, 1, 0), (3, 0): (1, 1, 1), (0, 1): (0, 0, 0), (1, 1): (1, 0, 0),
            (2, 1): (0, 1, 0), (3, 1): (0, 0, 1), (0, 2): (1, 1, 0), (1, 2): (0, 1, 1), (2, 2): (1, 0, 1), (3, 2): (h, h, h)}
    for (x, y), colour in want.items():
        assert_abs_diff_eq(c.pixel_at(x, y), colour, msg=f"pixel {x},{y}")


def test_ppm_parsing_ignores_comment_lines(lib):  # canvas.rs:369-384
    c = lib.canvas_from_ppm("""P3
        # this is a comment
        2 1
        # this, too
        255
        # another comment
        255 255 255
        # oh, no, comments in the pixel data!
        255 0 255
        """)
    assert_eq(c.pixel_at(0, 0), (1, 1, 1))
    assert_eq(c.pixel_at(1, 0), (1, 0, 1))


def test_ppm_parsing_allows_rgb_triplet_to_span_lines(lib):  # canvas.rs:386-398
    c = lib.canvas_from_ppm("P3\n        1 1\n        255\n        51\n        153\n        204\n        ")
    assert_eq(c.pixel_at(0, 0), (0.2, 0.6, 0.8))


def test_ppm_parsing_skips_empty_lines(lib):  # canvas.rs:400-416
    c = lib.canvas_from_ppm("\n        P3\n\n        1 1\n\n        255\n\n        51\n\n        153\n        204\n        ")
    assert_eq(c.pixel_at(0, 0), (0.2, 0.6, 0.8))


def test_ppm_parsing_respects_scale_setting(lib):  # canvas.rs:418-428
    c = lib.canvas_from_ppm("P3\n        2 2\n        100\n        100 100 100  50 50 50\n        75 50 25  0 0 0\n        ")
    assert_eq(c.pixel_at(0, 1), (0.75, 0.5, 0.25))


def test_too_much_pixel_data_is_an_error(lib):
    """The reference indexes data[height] and panics on the first extra pixel (canvas.rs:27-28,172); both
    restatements fail loudly instead."""
    from ray_tracer_challenge_b200.api import RtcError

    with pytest.raises(RtcError):
        lib.canvas_from_ppm("P3\n1 1\n255\n0 0 0 1 1 1\n")
    c = lib.canvas_from_ppm("P3\n2 1\n255\n9 9 9 7 7\n")  # an incomplete trailing triplet is dropped (canvas.rs:163)
    assert_eq(c.pixel_at(1, 0), (0, 0, 0))


UV_IMAGE_PPM = "P3\n10 10\n10\n" + "\n".join(
    "  ".join(" ".join([str((x + y) % 10)] * 3) for x in range(10)) for y in range(10)) + "\n"


def test_uv_mapping_an_image(oracle):  # uv.rs:640-672
    pattern = oracle.UVImage(oracle.canvas_from_ppm(UV_IMAGE_PPM))
    for u, v, want in ((0.0, 0.0, 0.9), (0.3, 0.0, 0.2), (0.6, 0.3, 0.1), (1.0, 1.0, 0.9)):
        assert_eq(oracle.probe.uv_color_at(pattern, u, v), (want, want, want), msg=f"uv ({u}, {v})")


def test_host_writer_and_reader_match_the_oracle(host, oracle):
    """Independent implementations (oracle: the reference's structure; host: table-driven single pass): identical
    bytes for awkward canvases, identical pixels for awkward files."""
    rng = np.random.default_rng(11)
    for w, h in ((1, 1), (5, 3), (23, 2), (24, 1), (70, 3), (7, 9), (0, 0), (3, 0)):
        data = rng.uniform(-0.3, 1.4, size=(h, w, 3)).astype(np.float32)
        if data.size > 6:
            flat = data.reshape(-1)
            flat[0], flat[1], flat[2], flat[3] = np.nan, np.inf, -np.inf, 0.5
            flat[4], flat[5] = np.float32(1.0) - np.float32(6e-8), 254.999 / 255.0
        a, b = host.new_canvas(w, h), oracle.new_canvas(w, h)
        a.data[:] = data
        b.data[:] = data
        assert a.to_ppm() == b.to_ppm(), (w, h)
    # widths whose rows end exactly at / just over the 70-column rule, values of 1-3 digits
    for w in range(1, 40):
        data = np.tile(rng.choice([0.0, 0.02, 0.5, 1.0], size=(1, w, 3)).astype(np.float32), (2, 1, 1))
        a, b = host.new_canvas(w, 2), oracle.new_canvas(w, 2)
        a.data[:] = data
        b.data[:] = data
        ppm = a.to_ppm()
        assert ppm == b.to_ppm(), w
        back_h, back_o = host.canvas_from_ppm(ppm), oracle.canvas_from_ppm(ppm)
        assert np.array_equal(back_h.data.view(np.uint32), back_o.data.view(np.uint32))
        assert back_h.to_ppm() == ppm  # write -> read -> write is the identity on 8-bit data
    from ray_tracer_challenge_b200 import scenes

    text = scenes.synthetic_ppm(37, 11, seed=3, scale=1000)
    assert np.array_equal(host.canvas_from_ppm(text).data.view(np.uint32), oracle.canvas_from_ppm(text).data.view(np.uint32))


def test_render_canvas_serialises_natively(oracle):
    """Camera::render's canvas goes through the same to_ppm (the demos' last step, e.g. soft_shadows.rs:61-62)."""
    from ray_tracer_challenge_b200 import scenes

    cam, world = scenes.default_world(oracle, 11, 11)
    canvas = cam.render(world, 5)
    lines = canvas.to_ppm().splitlines()
    assert lines[:3] == ["P3", "11 11", "255"]
    u8 = canvas.to_u8()
    assert [int(v) for v in " ".join(lines[3:]).split()] == [int(v) for v in u8.reshape(-1)]


def test_ppm_from_the_8_bit_plane_equals_to_ppm_of_the_f32_canvas():
    """Camera::render_b200_u8 ships only the 8-bit plane; its PPM must be the bytes Canvas::to_ppm (canvas.rs:58-96) writes
    for the f32 frame: same header, same 70-column wrap, same scale_color values (canvas.rs:39-43)."""
    import numpy as np

    import ray_tracer_challenge_b200 as rt

    host = rt.new_session()
    rng = np.random.default_rng(4)
    for w, h in ((5, 3), (10, 2), (23, 7), (1, 1)):
        data = rng.uniform(-0.2, 1.3, size=(h, w, 3)).astype(np.float32)
        canvas = host.new_canvas(w, h)
        canvas.data[...] = data
        want = canvas.to_ppm()
        u8 = np.clip(np.minimum(data * np.float32(255.0), np.float32(255.0)), 0.0, None).astype(np.uint8)  # scale_color
        got = rt.CanvasU8(w, h, u8, host).to_ppm()
        assert got == want, (w, h)
        assert all(len(line) <= 70 for line in got.splitlines())
