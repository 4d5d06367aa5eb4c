"""The GLSL vocabulary the pixel oracle stands on (oracle/ref_harness/glsl_shim.h + make_shader_cpp.py), checked against
an independent numpy evaluation of a probe shader written for this test (tests/glsl/probe.comp): swizzles, `out`
parameters, uniform / buffer blocks, literal suffixes, mix / mod / clamp / dot / length, texture() with clamp-to-edge,
repeat, nearest and linear filtering, imageStore's unorm conversion. CPU only; needs g++ (no /root/reference)."""
import os
import subprocess
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
HARNESS = os.path.join(ROOT, "oracle", "ref_harness")

f32 = np.float32
TEX = np.array([[[0.10, 0.20, 0.30, 1.00], [0.90, 0.10, 0.50, 0.25], [0.40, 0.80, 0.00, 0.75]],
                [[0.00, 1.00, 0.60, 0.50], [0.30, 0.30, 0.30, 0.30], [1.00, 0.00, 0.20, 0.10]]], f32)
WORDS = [0x10204080, 0xff00ff00, 0x01020304, 0x7f7f7f7f, 0xdeadbeef, 0x00000000, 0xffffffff, 0x80402010]


def tex_fetch(x, y, repeat_u):
    w, h = 3, 2
    x = x % w if repeat_u else min(max(x, 0), w - 1)
    y = min(max(y, 0), h - 1)
    return TEX[y, x]


def texture(u, v, linear, repeat_u):
    x, y = f32(u * f32(3)), f32(v * f32(2))
    if not linear:
        return tex_fetch(int(np.floor(x)), int(np.floor(y)), repeat_u)
    fx, fy = f32(x - f32(0.5)), f32(y - f32(0.5))
    x0, y0 = np.floor(fx), np.floor(fy)
    ax, ay = f32(fx - x0), f32(fy - y0)
    snap = f32(1.0 / 512.0)
    ax = f32(0) if ax < snap else (f32(1) if ax > f32(1) - snap else ax)
    ay = f32(0) if ay < snap else (f32(1) if ay > f32(1) - snap else ay)
    mix = lambda a, b, t: (a * (f32(1) - t) + b * t).astype(f32)
    x0, y0 = int(x0), int(y0)
    top = mix(tex_fetch(x0, y0, repeat_u), tex_fetch(x0 + 1, y0, repeat_u), ax)
    bot = mix(tex_fetch(x0, y0 + 1, repeat_u), tex_fetch(x0 + 1, y0 + 1, repeat_u), ax)
    return mix(top, bot, ay)


def expected(linear, repeat_u):
    out = np.zeros((2, 4, 4), np.uint8)
    scale = np.array([1.0, 0.5, 2.0, 0.75], f32)
    for y in range(2):
        for x in range(4):
            w = WORDS[x + 4 * y]
            v = (np.array([w & 0xff, (w >> 8) & 0xff, (w >> 16) & 0xff, w >> 24], f32) / f32(255)) * scale
            lo, hi = v[:2], v[2:]
            t = texture(f32((f32(x) + lo[1]) / f32(4)), f32((f32(y) + lo[0]) / f32(2)), linear, repeat_u)
            c = (np.array([hi[0], hi[1], t[3]], f32) + t[0]).astype(f32)
            c[:2] = c[:2] * c[1:3]          # c.rg *= c.gb
            c = (c * f32(0.5)).astype(f32)
            z = (c[::-1] + f32(0.25)).astype(f32)
            m = (z - f32(2) * np.floor(z / f32(2))).astype(f32)
            fo = np.clip(f32(1) - np.abs(f32(1) - m), 0, 1).astype(f32)
            length = f32(np.sqrt(f32(f32(lo[0] * lo[0]) + f32(lo[1] * lo[1]))))
            dot = f32(f32(f32(fo[0] * f32(0.125)) + f32(fo[1] * f32(0.125))) + f32(fo[2] * f32(0.125)))
            mixed = f32(f32(fo[0] * f32(0.75)) + f32(fo[1] * f32(0.25)))
            mval = f32(f32(mixed + dot) - min(length, f32(0.5)))
            px = np.array([fo[0], fo[1], mval, t[1]], f32)
            out[y, x] = np.rint(np.clip(px, 0, 1) * f32(255)).astype(np.uint8)
    return out


def expected_integers():
    """integerProbe() of tests/glsl/probe.comp in plain Python integers, over two runs of the 8 invocations."""
    def s16(v):
        return v - 0x10000 if v & 0x8000 else v

    counters, signed = [0] * 9, [-5, 9]
    for _ in range(2):
        for y in range(2):
            for x in range(4):
                w = WORDS[x + 4 * y]
                wi = w - (1 << 32) if w & 0x80000000 else w
                r = [s16(w & 0xffff), wi >> 16, 7, -3]
                t = [r[0] + 2, r[1] + 1]
                u = [int(max(np.float32(np.float32(v * 3) * np.float32(0.5) + np.float32(64.0)), np.float32(0))) for v in (t[0], t[1], t[0], t[1])]
                outside = [t[0] < r[2], t[1] < r[3], t[0] >= r[2], t[1] >= r[3]]
                slot = counters[0]
                counters[0] += 2
                before = signed[0]
                signed[0] = max(signed[0], t[0])
                swapped = signed[1]
                signed[1] = t[1]
                d = (u[2] - u[0]) & 0xffffffff
                val = (slot + (100 if any(outside) else 0) + (1000 if all(outside) else 0) + (before & 0xff) * 10000 +
                       (swapped & 0xf) * 1000000 + d + u[3]) & 0xffffffff
                counters[1 + x + 4 * y] = val
    return counters, signed


def test_probe_shader_through_the_shim(tmp_path):
    inc = tmp_path / "probe_comp.inc"
    with open(inc, "w") as fp:
        subprocess.check_call([sys.executable, os.path.join(HARNESS, "make_shader_cpp.py"),
                               os.path.join(HERE, "glsl", "probe.comp"), "probe_comp"], stdout=fp)
    exe = tmp_path / "probe"
    subprocess.check_call(["g++", "-std=c++17", "-O2", "-ffp-contract=off", "-w", "-I" + HARNESS, "-I" + str(tmp_path),
                           os.path.join(HERE, "glsl", "probe_main.cpp"), "-o", str(exe)])
    lines = subprocess.check_output([str(exe)], text=True).strip().split("\n")
    for variant, (linear, repeat_u) in enumerate(((True, False), (False, True))):
        got = np.array([int(v) for v in lines[variant].split()], np.uint8).reshape(2, 4, 4)
        assert np.array_equal(got, expected(linear, repeat_u)), (variant, got, expected(linear, repeat_u))
    # integer vectors / atomics: both variants ran, 8 invocations each, on the same counters
    got_counters = [int(v) for v in lines[2].split()]
    got_signed = [int(v) for v in lines[3].split()]
    want_counters, want_signed = expected_integers()
    assert got_counters == want_counters and got_signed == want_signed, (got_counters, want_counters, got_signed, want_signed)
    lines = lines[:2] + lines[4:]
    # 1e-7 off the centre of texel (1, 0): exactly that texel
    assert np.array_equal(np.array([float(v) for v in lines[2].split()], f32), TEX[0, 1])
