// rt_cull.h — host side: which pixels of the raster can see the scene at all (used by api.cu when it builds the pixel
// list of the beauty pass, and by the CPU harness in tests/hostsim so that the rule is testable without a GPU).
#pragma once
#include <algorithm>
#include <cmath>

#include "rt_scene.h"

namespace rt {

// 4x4 inverse in double (Gauss-Jordan with partial pivoting); false when singular.
inline bool invert4(const float* m, double* out) {
    double a[4][8];
    for (int r = 0; r < 4; r++)
        for (int c = 0; c < 4; c++) { a[r][c] = m[4 * r + c]; a[r][4 + c] = r == c ? 1.0 : 0.0; }
    for (int c = 0; c < 4; c++) {
        int piv = c;
        for (int r = c + 1; r < 4; r++) if (std::fabs(a[r][c]) > std::fabs(a[piv][c])) piv = r;
        if (std::fabs(a[piv][c]) < 1e-300) return false;
        if (piv != c) for (int k = 0; k < 8; k++) std::swap(a[piv][k], a[c][k]);
        const double inv = 1.0 / a[c][c];
        for (int k = 0; k < 8; k++) a[c][k] *= inv;
        for (int r = 0; r < 4; r++)
            if (r != c) { const double f = a[r][c]; if (f != 0.0) for (int k = 0; k < 8; k++) a[r][k] -= f * a[c][k]; }
    }
    for (int r = 0; r < 4; r++) for (int c = 0; c < 4; c++) out[4 * r + c] = a[r][4 + c];
    return true;
}

// Raster rectangle [x0, y0, x1, y1] (inclusive pixel coordinates) outside of which no camera ray can reach the scene bounds.
// A sample of pixel (px, py) leaves through raster point (x, y) = (px + dx, py + dy), dx, dy in [0, 1) (generate_ray,
// lib.rs:198-245), along the line camera_ray builds from raster_to_camera and camera_to_world (lib.rs:145-195):
//   pinhole       the line through camera_to_world * 0 with direction camera_to_world * (raster_to_camera * (x, y, 0));
//   orthographic  the line through camera_to_world * (raster_to_camera * (x, y, 0)) with direction camera_to_world * e_z.
// For each corner P of the (grown) bounds the raster point whose line passes through P is the solution of a 3x3 system in
// the columns of raster_to_camera (solved in double): the rectangle is the bounding box of the eight solutions plus two
// pixels. The map P -> (x, y) is projective; it keeps the box's image convex only while all corners lie on one side of the
// plane it sends to infinity, so the rectangle is used only when the eight determinants share a sign (box not straddling
// the camera plane). Thin-lens cameras (the origin moves over the aperture), non-affine matrices, scenes with an
// environment light (misses are lit) and empty scenes keep every pixel.
// Pixels outside the rectangle would have every sample dropped by raygen's per-sample bounds test anyway (their radiance is
// exactly 0: the reference's root-AABB reject, accel.rs:95); dropping them up front means they take no path slots, so a
// batch of the wavefront holds ~4x more live paths on frames like C3 (the box covers a quarter of the raster).
inline double det3(const double a[3], const double b[3], const double c[3]) {
    return a[0] * (b[1] * c[2] - b[2] * c[1]) - b[0] * (a[1] * c[2] - a[2] * c[1]) + c[0] * (a[1] * b[2] - a[2] * b[1]);
}
inline bool scene_raster_rect(const SceneD& sc, int rect[4]) {
    if (sc.env_texture != NONE || sc.camera.kind == 2u || sc.prim_count == 0) return false;
    const float* R = sc.camera.raster_to_camera.m;
    const float* C = sc.camera.camera_to_world.m;
    if (C[12] != 0.0f || C[13] != 0.0f || C[14] != 0.0f || C[15] != 1.0f) return false;   // camera_to_world must be affine
    const bool ortho = sc.camera.kind == 0u;
    if (ortho && (R[12] != 0.0f || R[13] != 0.0f || R[15] != 1.0f)) return false;       // apply_point must not divide
    double c2w_inv[16];
    if (!invert4(C, c2w_inv)) return false;
    const double a0[3] = {R[0], R[4], R[8]}, a1[3] = {R[1], R[5], R[9]}, a3[3] = {R[3], R[7], R[11]}, ez[3] = {0.0, 0.0, 1.0};
    double lo[2] = {1e300, 1e300}, hi[2] = {-1e300, -1e300}, sign0 = 0.0;
    for (int corner = 0; corner < 8; corner++) {
        const double p[3] = {(corner & 1) ? sc.bounds_hi[0] : sc.bounds_lo[0], (corner & 2) ? sc.bounds_hi[1] : sc.bounds_lo[1],
                             (corner & 4) ? sc.bounds_hi[2] : sc.bounds_lo[2]};
        double pc[3];
        for (int i = 0; i < 3; i++) pc[i] = c2w_inv[4 * i] * p[0] + c2w_inv[4 * i + 1] * p[1] + c2w_inv[4 * i + 2] * p[2] + c2w_inv[4 * i + 3];
        double third[3], rhs[3];
        if (ortho) for (int i = 0; i < 3; i++) { third[i] = ez[i]; rhs[i] = pc[i] - a3[i]; }       // x a0 + y a1 + t e_z = P_c - a3
        else for (int i = 0; i < 3; i++) { third[i] = -pc[i]; rhs[i] = -a3[i]; }                   // x a0 + y a1 + a3 = k P_c
        const double det = det3(a0, a1, third);
        const double scale = std::fabs(a0[0]) + std::fabs(a0[1]) + std::fabs(a0[2]) + std::fabs(a1[0]) + std::fabs(a1[1]) + std::fabs(a1[2]);
        const double mag = scale * scale * (std::fabs(third[0]) + std::fabs(third[1]) + std::fabs(third[2]));
        if (!(std::fabs(det) > 1e-9 * mag)) return false;
        const double sg = det > 0.0 ? 1.0 : -1.0;
        if (corner == 0) sign0 = sg;
        else if (sg != sign0) return false;
        const double v[2] = {det3(rhs, a1, third) / det, det3(a0, rhs, third) / det};
        for (int k = 0; k < 2; k++) {
            if (!std::isfinite(v[k])) return false;
            lo[k] = std::min(lo[k], v[k]);
            hi[k] = std::max(hi[k], v[k]);
        }
    }
    const double margin = 2.0;
    const double W = (double)sc.camera.width, H = (double)sc.camera.height;
    rect[0] = (int)std::floor(std::max(0.0, std::min(W, lo[0] - margin - 1.0)));   // pixel px covers [px, px + 1)
    rect[1] = (int)std::floor(std::max(0.0, std::min(H, lo[1] - margin - 1.0)));
    rect[2] = (int)std::ceil(std::max(-1.0, std::min(W - 1.0, hi[0] + margin)));
    rect[3] = (int)std::ceil(std::max(-1.0, std::min(H - 1.0, hi[1] + margin)));
    return true;
}

}  // namespace rt
