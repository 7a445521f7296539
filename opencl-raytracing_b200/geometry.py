"""f32 host-side math mirroring crates/raytracing/src/geometry/{matrix4x4,transform}.rs.

Only used while *describing* a scene (camera matrices, instance transforms); every per-ray use of these
matrices happens on the device. All arithmetic is numpy float32, accumulated in the reference's order
where it is cheap to do so (matmul: k-ascending dot, matrix4x4.rs:233-245).
"""
from __future__ import annotations

import math

import numpy as np

f32 = np.float32


def vec3(x, y, z) -> np.ndarray:
    return np.array([x, y, z], dtype=f32)


def _dot3(a, b) -> np.float32:
    return f32(f32(f32(a[0] * b[0]) + f32(a[1] * b[1])) + f32(a[2] * b[2]))


def length(v) -> np.float32:
    return f32(np.sqrt(_dot3(v, v)))


def unit(v) -> np.ndarray:
    # Vec3::unit = v * (1/len)  (vec3.rs:49-52,179-184)
    return (v * f32(f32(1.0) / length(v))).astype(f32)


def cross(u, v) -> np.ndarray:
    return np.array([f32(u[1] * v[2]) - f32(u[2] * v[1]), f32(u[2] * v[0]) - f32(u[0] * v[2]),
                     f32(u[0] * v[1]) - f32(u[1] * v[0])], dtype=f32)


def mat_identity() -> np.ndarray:
    return np.eye(4, dtype=f32)


def matmul(a: np.ndarray, b: np.ndarray) -> np.ndarray:
    """Matrix4x4::matmul (matrix4x4.rs:233-245): f32 dot accumulated k = 0..3."""
    m = np.zeros((4, 4), dtype=f32)
    for i in range(4):
        for j in range(4):
            d = f32(0.0)
            for k in range(4):
                d = f32(d + f32(a[i, k] * b[k, j]))
            m[i, j] = d
    return m


def mat_invert(m: np.ndarray) -> np.ndarray:
    """Matrix4x4::invert (matrix4x4.rs:72-198): adjugate / determinant in f32.

    The reference spells the 16 cofactors out term by term; here they come from 3x3 minors, which is
    the same polynomial (results can differ from the reference by an ulp in degenerate cases)."""
    m = m.astype(f32)

    def minor(r, c):
        rows = [i for i in range(4) if i != r]
        cols = [j for j in range(4) if j != c]
        s = m[np.ix_(rows, cols)]
        t0 = f32(s[0, 0] * f32(f32(s[1, 1] * s[2, 2]) - f32(s[1, 2] * s[2, 1])))
        t1 = f32(s[0, 1] * f32(f32(s[1, 0] * s[2, 2]) - f32(s[1, 2] * s[2, 0])))
        t2 = f32(s[0, 2] * f32(f32(s[1, 0] * s[2, 1]) - f32(s[1, 1] * s[2, 0])))
        return f32(f32(t0 - t1) + t2)

    cof = np.zeros((4, 4), dtype=f32)
    for r in range(4):
        for c in range(4):
            sign = f32(1.0) if (r + c) % 2 == 0 else f32(-1.0)
            cof[r, c] = f32(sign * minor(r, c))
    det = f32(0.0)
    for c in range(4):
        det = f32(det + f32(m[0, c] * cof[0, c]))
    if det == 0.0:
        raise ValueError("failed to invert matrix")
    inv_det = f32(f32(1.0) / det)
    return (cof.T * inv_det).astype(f32)


class Transform:
    """crates/raytracing/src/geometry/transform.rs:3-83 — a matrix and its inverse, row-major."""

    __slots__ = ("forward", "inverse")

    def __init__(self, forward: np.ndarray, inverse: np.ndarray | None = None):
        self.forward = np.asarray(forward, dtype=f32).reshape(4, 4)
        self.inverse = mat_invert(self.forward) if inverse is None else np.asarray(inverse, dtype=f32).reshape(4, 4)

    @staticmethod
    def identity() -> "Transform":
        return Transform(mat_identity(), mat_identity())

    @staticmethod
    def translate(d) -> "Transform":
        f, i = mat_identity(), mat_identity()
        f[:3, 3] = np.asarray(d, dtype=f32)
        i[:3, 3] = -np.asarray(d, dtype=f32)
        return Transform(f, i)

    @staticmethod
    def scale(s) -> "Transform":
        s = np.asarray(s, dtype=f32)
        f, i = mat_identity(), mat_identity()
        for k in range(3):
            f[k, k] = s[k]
            i[k, k] = f32(f32(1.0) / s[k])
        return Transform(f, i)

    @staticmethod
    def rotate(theta: float, v) -> "Transform":
        """Matrix4x4::rotation (matrix4x4.rs:265-313), inverse = transpose (transform.rs:26-33)."""
        v = np.asarray(v, dtype=f32)
        ct, st = f32(math.cos(f32(theta))), f32(math.sin(f32(theta)))

        def rot(u):
            v_c = (v * _dot3(u, v)).astype(f32)
            v1 = (u - v).astype(f32)
            v2 = cross(v, v1)
            return (v_c + v1 * ct + v2 * st).astype(f32)

        f = mat_identity()
        f[:3, 0] = rot(vec3(1, 0, 0))
        f[:3, 1] = rot(vec3(0, 1, 0))
        f[:3, 2] = rot(vec3(0, 0, 1))
        return Transform(f, f.T.copy())

    def compose(self, other: "Transform") -> "Transform":
        """self first, then other (transform.rs:42-49)."""
        return Transform(matmul(other.forward, self.forward), matmul(self.inverse, other.inverse))

    def invert(self) -> "Transform":
        return Transform(self.inverse.copy(), self.forward.copy())

    @staticmethod
    def from_matrix(m) -> "Transform":
        return Transform(np.asarray(m, dtype=f32).reshape(4, 4))

    @staticmethod
    def look_at(camera_pos, target_pos, up, swap_handedness: bool) -> "Transform":
        """transform.rs:96-149: +z forward, x = -unit(view x up), y = view x camera_x."""
        camera_pos = np.asarray(camera_pos, dtype=f32)
        view = unit((np.asarray(target_pos, dtype=f32) - camera_pos).astype(f32))
        cx = (-unit(cross(view, np.asarray(up, dtype=f32)))).astype(f32)
        cy = cross(view, cx)
        if swap_handedness:
            cx = (-cx).astype(f32)
        m = mat_identity()
        m[:3, 0], m[:3, 1], m[:3, 2], m[:3, 3] = cx, cy, view, camera_pos
        return Transform(m)

    def apply_point(self, p) -> np.ndarray:
        p = np.asarray(p, dtype=f32)
        h = np.array([p[0], p[1], p[2], 1.0], dtype=f32)
        r = self.forward @ h
        return (r[:3] / r[3]).astype(f32)
