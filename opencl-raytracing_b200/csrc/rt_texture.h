// rt_texture.h — texture fetch in `shade` and the mip-pyramid build kernels' per-thread bodies.
//
// Replaces CpuTextures::{new,sample,point_sample,bilerp_sample,mip_level,sample_image_texture,
// texture_mip_level} and mipmap::generate_mips (crates/raytracing-cpu/src/texture.rs:114-165, 207-480),
// WrapMode::apply (crates/raytracing/src/materials/texture.rs:45-68) and Image::get_pixel
// (crates/raytracing/src/materials/image.rs:56-121). Software filtering on purpose: the reference's
// filter arithmetic (clamped floor/ceil bilinear, un-clamped fract(level) trilinear, bilinear fallback on
// the ORIGINAL image) is not what the texture units compute.
#pragma once
#include "rt_scene.h"

namespace rt {

// materials.rs:702-796
struct MatCtx { V2 uv; float dudx, dudy, dvdx, dvdy; };
RT_HD MatCtx matctx_no_aa(V2 uv) { MatCtx c; c.uv = uv; c.dudx = c.dudy = c.dvdx = c.dvdy = 0.0f; return c; }

RT_HD float image_channel(const SceneD& sc, const ImageD& im, uint32_t x, uint32_t y, uint32_t c) {
    if (c >= im.channels) return 0.0f;
    size_t idx = ((size_t)y * im.width + x) * im.channels + c;
    const uint8_t* base = sc.image_bytes + im.byte_offset;
    if (im.format == 0) return (float)ldg(base + idx) / 255.0f;
    if (im.format == 1) return (float)ldg((const uint16_t*)base + idx) / 65535.0f;
    return ldg((const float*)base + idx) / 1.0f;
}
RT_HD V4 image_pixel(const SceneD& sc, const ImageD& im, uint32_t x, uint32_t y) {
    return mk4(image_channel(sc, im, x, y, 0), image_channel(sc, im, x, y, 1), image_channel(sc, im, x, y, 2),
               image_channel(sc, im, x, y, 3));
}

RT_HD float wrap_apply(uint32_t mode, float x) {  // texture.rs:45-68
    if (mode == 0) { float f = rs_fract(x); return f < 0.0f ? 1.0f + f : f; }
    if (mode == 1) {
        float f = rs_fract(x);
        float rep = f < 0.0f ? 1.0f + f : f;
        int32_t fl = rs_as_i32(floorf(x));
        int32_t r = fl % 2;
        if (r < 0) r += 2;
        return r == 1 ? 1.0f - rep : rep;
    }
    return rs_clamp(x, 0.0f, 1.0f);
}

RT_HD V4 point_sample(const SceneD& sc, const ImageD& im, float u, float v) {  // texture.rs:235-246
    float w = (float)im.width, h = (float)im.height;
    float x = u * w - 0.5f, y = v * h - 0.5f;
    uint32_t xi = rs_as_u32(rs_clamp(roundf(x), 0.0f, w - 1.0f));
    uint32_t yi = rs_as_u32(rs_clamp(roundf(y), 0.0f, h - 1.0f));
    return image_pixel(sc, im, xi, yi);
}
RT_HD V4 bilerp_sample(const SceneD& sc, const ImageD& im, float u, float v) {  // texture.rs:248-268
    float w = (float)im.width, h = (float)im.height;
    float x = u * w - 0.5f, y = v * h - 0.5f;
    uint32_t x0 = rs_as_u32(rs_clamp(floorf(x), 0.0f, w - 1.0f));
    uint32_t x1 = rs_as_u32(rs_clamp(ceilf(x), 0.0f, w - 1.0f));
    uint32_t y0 = rs_as_u32(rs_clamp(floorf(y), 0.0f, h - 1.0f));
    uint32_t y1 = rs_as_u32(rs_clamp(ceilf(y), 0.0f, h - 1.0f));
    float xf = rs_clamp(rs_fract(x), 0.0f, 1.0f), yf = rs_clamp(rs_fract(y), 0.0f, 1.0f);
    V4 p00 = image_pixel(sc, im, x0, y0), p01 = image_pixel(sc, im, x1, y0);
    V4 p10 = image_pixel(sc, im, x0, y1), p11 = image_pixel(sc, im, x1, y1);
    V4 u0 = p00 * (1.0f - xf) + p01 * xf;
    V4 u1 = p10 * (1.0f - xf) + p11 * xf;
    return u0 * (1.0f - yf) + u1 * yf;
}
RT_HD bool mip_level_of(uint32_t mip0_width, const MatCtx& c, float& level) {  // texture.rs:270-297
    float dx = sqrtf(c.dudx * c.dudx + c.dvdx * c.dvdx);
    float dy = sqrtf(c.dudy * c.dudy + c.dvdy * c.dvdy);
    float larger = fmaxf(dx, dy);
    if (larger <= 0.0f) return false;
    float half_pixel = 1.0f / (2.0f * (float)mip0_width);
    level = log2f(larger / half_pixel);
    return true;
}

RT_HD_CALL V4 sample_image_texture(const SceneD& sc, const TextureD& tx, const MatCtx& c) {  // texture.rs:299-357
    const ImageD& im = sc.images[tx.image];
    float u = wrap_apply(tx.wrap, c.uv.x), v = wrap_apply(tx.wrap, c.uv.y);
    if (tx.filter == 0) return point_sample(sc, im, u, v);
    if (tx.filter == 1 || tx.mip_base == NONE) return bilerp_sample(sc, im, u, v);
    const MipChain mc = sc.mips[tx.mip_base];
    const ImageD& mip0 = sc.images[mc.first_image];
    float level;
    if (!mip_level_of(mip0.width, c, level)) return bilerp_sample(sc, im, u, v);
    float maxl = (float)(mc.level_count - 1);
    uint32_t lower = rs_as_u32(floorf(rs_clamp(level, 0.0f, maxl)));
    uint32_t upper = rs_as_u32(ceilf(rs_clamp(level, 0.0f, maxl)));
    float t = rs_fract(level);
    V4 a = bilerp_sample(sc, sc.images[mc.first_image + lower], u, v);
    V4 b = bilerp_sample(sc, sc.images[mc.first_image + upper], u, v);
    return t * b + (1.0f - t) * a;
}

RT_HD_CALL V4 checker_sample(const TextureD& tx, const MatCtx& c) {  // texture.rs:376-434
    V4 color1 = mk4(tx.value[0], tx.value[1], tx.value[2], tx.value[3]);
    V4 color2 = mk4(tx.value2[0], tx.value2[1], tx.value2[2], tx.value2[3]);
    float u = c.uv.x - floorf(c.uv.x), v = c.uv.y - floorf(c.uv.y);
    if ((c.dudx == 0.0f && c.dvdx == 0.0f) || (c.dudy == 0.0f && c.dvdy == 0.0f))
        return ((u > 0.5f) != (v > 0.5f)) ? color1 : color2;
    float srx = sqrtf(c.dudx * c.dudx + c.dvdx * c.dvdx);
    float sry = sqrtf(c.dudy * c.dudy + c.dvdy * c.dvdy);
    float sigma = 0.1f * fmaxf(srx, sry);
    float a = u < 0.25f ? u : (u < 0.75f ? -(u - 0.5f) : u - 1.0f);
    float b = v < 0.25f ? v : (v < 0.75f ? -(v - 0.5f) : v - 1.0f);
    float xz = a / (sqrtf(2.0f) * sigma), yz = b / (sqrtf(2.0f) * sigma);
    float xf = 0.5f * (1.0f + erff(xz)), yf = 0.5f * (1.0f + erff(yz));
    xf = v > 0.5f ? xf : 1.0f - xf;
    yf = u > 0.5f ? yf : 1.0f - yf;
    float factor = xf * yf;
    return factor * color1 + (1.0f - factor) * color2;
}

// texture.rs:359-459. Scale / Mix nest; the nesting depth is bounded at compile time (the importers
// produce depth <= 1: Scale(Image, Constant), scene.rs:344-356).
template <int DEPTH>
RT_HD V4 sample_texture(const SceneD& sc, uint32_t tex_id, const MatCtx& c) {
    const TextureD& tx = sc.textures[tex_id];
    switch (tx.kind) {
        case 0: return sample_image_texture(sc, tx, c);
        case 1: return mk4(tx.value[0], tx.value[1], tx.value[2], tx.value[3]);
        case 2: return checker_sample(tx, c);
        case 3:
            if constexpr (DEPTH > 0) return sample_texture<DEPTH - 1>(sc, tx.a, c) * sample_texture<DEPTH - 1>(sc, tx.b, c);
            else return mk4(0, 0, 0, 0);
        default:
            if constexpr (DEPTH > 0) {
                V4 one = mk4(1, 1, 1, 1), zero = mk4(0, 0, 0, 0);
                V4 cv = sample_texture<DEPTH - 1>(sc, tx.c, c);
                V4 bv = (cv == zero) ? zero : sample_texture<DEPTH - 1>(sc, tx.b, c);
                V4 av = (cv == one) ? zero : sample_texture<DEPTH - 1>(sc, tx.a, c);
                return (one - cv) * av + cv * bv;
            } else return mk4(0, 0, 0, 0);
    }
}
constexpr int TEXTURE_NEST = 3;
RT_HD_CALL V4 tex_general(const SceneD& sc, uint32_t tex_id, const MatCtx& c) { return sample_texture<TEXTURE_NEST>(sc, tex_id, c); }
RT_HD V4 tex(const SceneD& sc, uint32_t tex_id, const MatCtx& c) {
    const TextureD& tx = sc.textures[tex_id];
    if (tx.kind == 1) return mk4(ldg(&tx.value[0]), ldg(&tx.value[1]), ldg(&tx.value[2]), ldg(&tx.value[3]));  // ConstantTexture inline
    return tex_general(sc, tex_id, c);
}

RT_HD bool texture_mip_level(const SceneD& sc, uint32_t tex_id, const MatCtx& c, float& level) {  // texture.rs:461-480
    const TextureD& tx = sc.textures[tex_id];
    if (tx.kind == 0 && tx.filter == 2 && tx.mip_base != NONE) return mip_level_of(sc.images[sc.mips[tx.mip_base].first_image].width, c, level);
    return false;
}

// ---- mip pyramid build (image 0.25.8 imageops::resize(.., Lanczos3): vertical pass unclamped, horizontal
// pass clamped to [0,1]; then texture.rs:86-111 cast back to the source sample type) ------------------
RT_HD float lanczos3(float x) {
    if (!(fabsf(x) < 3.0f)) return 0.0f;
    float a = x * PI, b = (x / 3.0f) * PI;
    float s1 = x == 0.0f ? 1.0f : sinf(a) / a;
    float s2 = (x / 3.0f) == 0.0f ? 1.0f : sinf(b) / b;
    return s1 * s2;
}

// One output sample of a 1-D Lanczos3 resample along `axis` (0 = vertical: src [h][w][ch] -> dst [nh][w][ch];
// 1 = horizontal: src [h][w][ch] -> dst [h][nw][ch], clamped).
RT_HD void resize_body(uint32_t idx, const float* src, float* dst, uint32_t w, uint32_t h, uint32_t ch, uint32_t n_out, int axis) {
    uint32_t c = idx % ch;
    uint32_t rest = idx / ch;
    uint32_t ox, oy, in_len;
    if (axis == 0) { ox = rest % w; oy = rest / w; in_len = h; }
    else { ox = rest % n_out; oy = rest / n_out; in_len = w; }
    uint32_t o = axis == 0 ? oy : ox;
    float ratio = (float)in_len / (float)n_out;
    float sratio = ratio < 1.0f ? 1.0f : ratio;
    float support = 3.0f * sratio;
    float in = ((float)o + 0.5f) * ratio;
    long long left = (long long)floorf(in - support);
    if (left < 0) left = 0;
    if (left > (long long)in_len - 1) left = (long long)in_len - 1;
    long long right = (long long)ceilf(in + support);
    if (right < left + 1) right = left + 1;
    if (right > (long long)in_len) right = (long long)in_len;
    in = in - 0.5f;
    float sum = 0.0f;
    for (long long i = left; i < right; i++) sum += lanczos3(((float)i - in) / sratio);
    float t = 0.0f;
    for (long long i = left; i < right; i++) {
        float wgt = lanczos3(((float)i - in) / sratio) / sum;
        size_t si = axis == 0 ? ((size_t)i * w + ox) * ch + c : ((size_t)oy * w + (size_t)i) * ch + c;
        t += src[si] * wgt;
    }
    if (axis == 0) dst[((size_t)oy * w + ox) * ch + c] = t;
    else dst[((size_t)oy * n_out + ox) * ch + c] = rs_clamp(t, 0.0f, 1.0f);
}

RT_HD void to_f32_body(uint32_t idx, const uint8_t* src, uint32_t format, float* dst) {
    if (format == 0) dst[idx] = (float)src[idx] / 255.0f;
    else if (format == 1) dst[idx] = (float)((const uint16_t*)src)[idx] / 65535.0f;
    else dst[idx] = ((const float*)src)[idx];
}
RT_HD void cast_body(uint32_t idx, const float* src, uint32_t format, uint8_t* dst) {  // texture.rs:86-111
    if (format == 0) dst[idx] = (uint8_t)roundf(rs_clamp(src[idx], 0.0f, 1.0f) * 255.0f);
    else if (format == 1) ((uint16_t*)dst)[idx] = (uint16_t)roundf(rs_clamp(src[idx], 0.0f, 1.0f) * 65535.0f);
    else ((float*)dst)[idx] = src[idx];
}

}  // namespace rt
