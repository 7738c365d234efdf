"""Conformance of tests/refexec (the interpreters that execute the reference's own JS / GLSL text
to mint tests/golden/ref_pins.*) with the language rules the reference's code relies on.

Every expectation below is a fact of ECMAScript or GLSL ES 3.00 / OpenGL ES 3.0, written out by
hand -- none comes from the oracle or the CUDA library -- so a shared misreading between the
interpreters and oracle/lbm_ref.c would show up here (round-1 verdict, "What's weak" #3):
operator precedence and associativity, `**`, integer division and int() truncation, NEAREST texel
addressing with CLAMP_TO_EDGE at the vUV extremes, mix(), fp32 rounding per operation, literal
rounding, RGBA8 conversion (what the page's readPixels sees).
"""
import math

import numpy as np
import pytest

from refexec.glslrun import Sampler, Shader, Vec, f32, lit, run_pass
from refexec.jsrun import Interp


def js(expr, pre=""):
    it = Interp()
    it.run(pre + f"\nfunction __t() {{ return {expr}; }}")
    return it.call("__t")


@pytest.mark.parametrize("expr,want", [
    ("2 + 3 * 4", 14.0), ("(2 + 3) * 4", 20.0), ("2 - 3 - 4", -5.0), ("24 / 4 / 2", 3.0),
    ("2 ** 3 ** 2", 512.0),                      # ** is right associative
    ("2 * 3 ** 2", 18.0),                        # and binds tighter than *
    ("-(2 ** 2)", -4.0),
    ("7 % 3", 1.0), ("-7 % 3", -1.0),            # remainder takes the sign of the dividend
    ("1 / 3", 1.0 / 3.0), ("7 / 2", 3.5),        # no integer division in JS
    ("1 + 2 < 4 && 3 > 2", True), ("1 < 2 === true", True),
    ("(1 > 2) !== (3 > 2)", True),               # the mask rasteriser's crossing test, HTML:168
    ("0.1 + 0.2", 0.1 + 0.2), ("0.1 * 3", 0.30000000000000004),
    ("1e-12 + 1", 1.000000000001),
    ("Math.hypot(3, 4)", 5.0), ("Math.max(1, 5, 3)", 5.0), ("Math.min(2, -1)", -1.0),
    ("Math.floor(-0.5)", -1.0), ("Math.ceil(-0.5)", 0.0), ("Math.ceil(3.0000001)", 4.0),
    ("Math.round(2.5)", 3.0), ("Math.round(-2.5)", -2.0), ("Math.round(0.49999999999999994)", 0.0),
    ("Math.abs(-3)", 3.0), ("Math.sqrt(2)", math.sqrt(2.0)), ("Math.PI", math.pi),
    ("1 / 0", math.inf), ("isFinite(1 / 0)", False), ("isFinite(0 / 0)", False),
    ("true ? 1 : 2", 1.0), ("0 ? 1 : 2", 2.0),
])
def test_js_expressions(expr, want):
    got = js(expr)
    if isinstance(want, bool):
        assert got is want or got == want
    else:
        assert float(got) == want and math.copysign(1, float(got)) == math.copysign(1, want)


def test_js_nan_and_typed_arrays():
    assert math.isnan(js("0 / 0"))
    assert js("NaN === NaN") is False
    # Float32Array rounds on store (equilibriumInitData, HTML:474-490), Uint8Array holds 0..255
    got = js("a[0]", "const a = new Float32Array(2); a[0] = 0.1;")
    assert got == float(np.float32(0.1)) and got != 0.1
    assert js("m[1]", "const m = new Uint8Array(3); m[1] = 255;") == 255
    # array sort with a comparator sorts numerically (HTML:173 `xs.sort((a,b)=>a-b)`)
    assert list(js("xs", "const xs = [10, 9, 1, 100]; xs.sort((a, b) => a - b);")) == [1, 9, 10, 100]
    # for loops, compound assignment, destructuring as the geometry code uses them
    assert js("s", "let s = 0; for (let i = 0; i < 5; i++) { s += i * 2; }") == 20
    assert js("x + y", "const [x, y] = [3, 4];") == 7


def glsl_value(expr, decl="float", pre=""):
    sh = Shader("precision highp float;\nout vec4 o;\n" + pre +
                f"\nvoid main() {{ {decl} t = {expr}; o = vec4(float(t), 0.0, 0.0, 1.0); }}")
    return sh.run_fragment({})["o"].v[0]


@pytest.mark.parametrize("expr,decl,want", [
    ("2.0 + 3.0 * 4.0", "float", 14.0), ("2.0 - 3.0 - 4.0", "float", -5.0), ("24.0 / 4.0 / 2.0", "float", 3.0),
    ("7 / 2", "int", 3), ("-7 / 2", "int", -3),             # integer division truncates toward zero
    ("int(2.9)", "int", 2), ("int(-2.9)", "int", -2),       # and so does int()
    ("1.0 / 3.0", "float", float(np.float32(1) / np.float32(3))),
    ("4.0 / 9.0", "float", float(np.float32(4) / np.float32(9))),      # the shader's w0, HTML:234
    ("1.0 / 36.0", "float", float(np.float32(1) / np.float32(36))),
    ("0.1 + 0.2", "float", float(np.float32(0.1) + np.float32(0.2))),  # fp32 literals, fp32 sum
    ("clamp(2.5, 0.5, 2.0)", "float", 2.0), ("clamp(0.1, 0.5, 2.0)", "float", 0.5),
    ("max(1.0, 2.0)", "float", 2.0), ("min(1.0, 2.0)", "float", 1.0),
    ("sqrt(2.0)", "float", float(np.sqrt(np.float32(2)))),
    ("floor(-0.5)", "float", -1.0),
    ("length(vec2(3.0, 4.0))", "float", 5.0),
])
def test_glsl_expressions(expr, decl, want):
    assert float(glsl_value(expr, decl)) == want


def test_glsl_every_operation_rounds_to_fp32():
    # (a*b)+c with an exactly representable product only in float64: fp32 must round the product first
    a, b, c = np.float32(1.0000004), np.float32(1.0000004), np.float32(-1.0)
    want = float(np.float32(a * b) + c)
    fused = float(np.float32(np.float64(a) * np.float64(b) + np.float64(c)))
    assert want != fused                      # the case distinguishes separate rounding from an FMA
    got = glsl_value("a * b + c", pre="const float a = 1.0000004; const float b = 1.0000004; const float c = -1.0;")
    assert float(got) == want
    # literals are rounded once, decimal -> fp32 (0.35, 0.58 ... are not float64 values cast down twice)
    assert float(lit("0.58")) == float(np.float32(0.58)) and float(lit("16777217.0")) == 16777216.0


def test_glsl_mix_and_swizzles():
    sh = Shader("precision highp float;\nout vec4 o;\n"
                "void main() { vec3 a = vec3(0.0, 1.0, 2.0); vec3 b = vec3(4.0, 5.0, 6.0);"
                " vec3 m = mix(a, b, 0.25); o = vec4(m.z, m.y, m.x, 1.0); }")
    o = sh.run_fragment({})["o"].v
    # mix(x, y, a) = x*(1-a) + y*a (GLSL ES 3.00 section 8.3)
    assert [float(v) for v in o] == [3.0, 2.0, 1.0, 1.0]


def test_texture_nearest_clamp_to_edge_addressing():
    """texture() of a NEAREST / CLAMP_TO_EDGE sampler (HTML:438-458): texel = floor(uv * size), clamped.
    The step shader samples at vUV +- k*texel, i.e. exactly at texel centres; at the lattice border
    the tap lands outside [0,1] and must return the edge texel."""
    w, h = 4, 3
    arr = np.arange(h * w * 4, dtype=np.float32).reshape(h, w, 4)
    s = Sampler(arr)

    def at(u, v):
        return [float(c) for c in s.fetch(Vec([f32(u), f32(v)])).v]

    for y in range(h):
        for x in range(w):
            assert at((x + 0.5) / w, (y + 0.5) / h) == list(arr[y, x])          # texel centres
    assert at(-0.125, 0.5 / h) == list(arr[0, 0])                               # left of the lattice -> edge
    assert at(1.125, 0.5 / h) == list(arr[0, w - 1])
    assert at(0.5 / w, -0.2) == list(arr[0, 0]) and at(0.5 / w, 1.2) == list(arr[h - 1, 0])
    assert at(0.25, 0.0) == list(arr[0, 1])                                     # texel boundary belongs to the upper texel
    assert at(1.0, 1.0) == list(arr[h - 1, w - 1])
    # single-channel (R8 mask) textures read as (r, 0, 0, 1)
    m = Sampler(np.array([[[1.0], [0.0]]], np.float32))
    assert [float(c) for c in m.fetch(Vec([f32(0.25), f32(0.5)])).v] == [1.0, 0.0, 0.0, 1.0]


def test_run_pass_draws_one_fragment_per_pixel_at_texel_centres():
    sh = Shader("precision highp float;\nin vec2 vUV;\nout vec4 o;\n"
                "void main() { o = vec4(vUV.x, vUV.y, 0.0, 1.0); }")
    out = run_pass(sh, 4, 2, {})["o"]
    assert out.shape == (2, 4, 4)
    assert np.array_equal(out[:, :, 0], np.tile(np.float32([0.125, 0.375, 0.625, 0.875]), (2, 1)))
    assert np.array_equal(out[:, :, 1], np.float32([[0.25] * 4, [0.75] * 4]))


def test_rgba8_conversion_rounds_to_nearest():
    """The framebuffer conversion of RENDER_FS_SRC's output: clamp to [0,1], scale by 255, round to
    nearest (OpenGL ES 3.0 section 2.1.6.1).  make_ref_pins.py and the oracle must agree with it."""
    from golden.make_ref_pins import to_unorm8
    vals = np.float32([-0.5, 0.0, 0.5 / 255, 0.4999 / 255, 1.0 / 255, 0.5, 254.5 / 255, 1.0, 7.0])
    assert list(to_unorm8(vals)) == [0, 0, 1, 0, 1, 128, 255, 255, 255]
