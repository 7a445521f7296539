// rt_bsdf.h — material evaluation in `shade`.
//
// Replaces CpuBsdf::{evaluate_bsdf,evaluate_pdf,sample_bsdf,is_delta_bsdf,components}, refract,
// fresnel_dielectric, fresnel_complex, microfacet::*, phase_function::* and CpuMaterial::{get_bsdf,
// get_mip_level,get_albedo} (crates/raytracing-cpu/src/materials.rs:10-1541), Complex::sqrt
// (crates/raytracing/src/geometry/complex.rs:197-216). The reference's boxed, recursive BSDF enum becomes
// two flat value types: `Bsdf` (one interface) and `Layered` (top + bottom interface + medium), so a
// thread keeps everything in registers and the LayeredBxDF random walk needs no recursion.
#pragma once
#include "rt_sampler.h"
#include "rt_texture.h"

namespace rt {

enum : uint32_t {  // materials.rs:90-103
    NONSPEC_REFL = 1, SPEC_REFL = 2, NONSPEC_TRANS = 4, SPEC_TRANS = 8,
    REFLECTION = NONSPEC_REFL | SPEC_REFL, TRANSMISSION = NONSPEC_TRANS | SPEC_TRANS,
    SPECULAR = SPEC_REFL | SPEC_TRANS, NONSPECULAR = NONSPEC_REFL | NONSPEC_TRANS, ALL_COMPONENTS = 15
};
enum : uint32_t { B_DIFFUSE = 0, B_SMOOTH_DIELECTRIC, B_SMOOTH_CONDUCTOR, B_ROUGH_CONDUCTOR, B_ROUGH_DIELECTRIC, B_LAYERED };
enum : int { S_VALID = 0, S_NULL = 1, S_INVALID = 2 };

struct BsdfSample { V3 wi, f; float pdf; uint32_t component; };

RT_HD int validate_sample(const BsdfSample& s) {  // materials.rs:28-39
    bool bad = !finite_f(s.f.x) || !finite_f(s.f.y) || !finite_f(s.f.z) || !finite_f(s.pdf) || !finite_f(s.wi.x) ||
               !finite_f(s.wi.y) || !finite_f(s.wi.z);
    if (bad || s.pdf <= 0.0f || popc32(s.component) != 1) return S_INVALID;
    return S_VALID;
}

struct Cx { float re, im; };
RT_HD Cx cx(float re, float im) { Cx c; c.re = re; c.im = im; return c; }
RT_HD Cx operator*(Cx a, Cx b) { return cx(a.re * b.re - a.im * b.im, a.im * b.re + a.re * b.im); }
RT_HD Cx operator*(Cx a, float s) { return cx(a.re * s, a.im * s); }
RT_HD Cx operator+(Cx a, Cx b) { return cx(a.re + b.re, a.im + b.im); }
RT_HD Cx operator-(Cx a, Cx b) { return cx(a.re - b.re, a.im - b.im); }
RT_HD Cx operator+(Cx a, float s) { return cx(a.re + s, a.im); }
RT_HD Cx operator-(Cx a) { return cx(-a.re, -a.im); }
RT_HD Cx operator/(Cx a, Cx b) {
    float den = b.re * b.re + b.im * b.im;
    return cx((a.re * b.re + a.im * b.im) / den, (a.im * b.re - a.re * b.im) / den);
}
RT_HD float csqmag(Cx a) { return a.re * a.re + a.im * a.im; }
RT_HD Cx csqrt(Cx a) {  // polar form
    float r = sqrtf(csqmag(a)), theta = atan2f(a.im, a.re);
    float sr = sqrtf(r), ht = theta / 2.0f;
    return cx(sr * cosf(ht), sr * sinf(ht));
}

RT_HD bool refract(float eta, V3 wo, V3 normal, V3& out) {  // materials.rs:992-1009
    float cos_i = dot(wo, normal);
    if (cos_i < 0.0f) { eta = 1.0f / eta; cos_i = -cos_i; normal = -normal; }
    float sin2_i = 1.0f - cos_i * cos_i;
    float sin2_t = sin2_i / (eta * eta);
    if (sin2_t >= 1.0f) return false;
    float cos_t = sqrtf(1.0f - sin2_t);
    out = -wo / eta + (cos_i / eta - cos_t) * normal;
    return true;
}
RT_HD float fresnel_dielectric(float cos_i, float eta) {  // materials.rs:1018-1041
    if (cos_i < 0.0f) { eta = 1.0f / eta; cos_i = -cos_i; }
    float sin2_i = 1.0f - cos_i * cos_i;
    float sin2_t = sin2_i / (eta * eta);
    if (sin2_t >= 1.0f) return 1.0f;
    float cos_t = sqrtf(1.0f - sin2_t);
    float r_parl = (eta * cos_i - cos_t) / (eta * cos_i + cos_t);
    float r_perp = (cos_i - eta * cos_t) / (cos_i + eta * cos_t);
    return (r_parl * r_parl + r_perp * r_perp) / 2.0f;
}
RT_HD_CALL float fresnel_complex(float cos_i, Cx eta) {  // materials.rs:1045-1065
    float sin2_i = 1.0f - cos_i * cos_i;
    Cx sin2_t = cx(sin2_i, 0.0f) / (eta * eta);
    Cx cos2_t = -sin2_t + 1.0f;
    Cx cos_t = csqrt(cos2_t);
    Cx r_parl = (eta * cos_i - cos_t) / (eta * cos_i + cos_t);
    Cx r_perp = (cx(cos_i, 0.0f) - eta * cos_t) / (cx(cos_i, 0.0f) + eta * cos_t);
    return (csqmag(r_parl) + csqmag(r_perp)) / 2.0f;
}
RT_HD V3 fresnel_complex3(float cos_i, V3 eta, V3 kappa) {
    return mk3(fresnel_complex(cos_i, cx(eta.x, kappa.x)), fresnel_complex(cos_i, cx(eta.y, kappa.y)),
               fresnel_complex(cos_i, cx(eta.z, kappa.z)));
}

// ---- microfacet (Trowbridge-Reitz), materials.rs:1068-1474 -----------------------------------------
RT_HD float mf_distribution(V3 wm, float ax, float ay) {
    float c2 = wm.z * wm.z, s2 = 1.0f - c2;
    float e = (wm.x * wm.x) / (ax * ax) + (wm.y * wm.y) / (ay * ay);
    float t = (1.0f + (s2 / c2) * e) * (1.0f + (s2 / c2) * e);
    return 1.0f / (PI * ax * ay * c2 * c2 * t);
}
RT_HD float mf_lambda(V3 w, float ax, float ay) {
    float c2 = w.z * w.z, s2 = 1.0f - c2, tan2 = s2 / c2;
    float a2 = ax * ax * w.x * w.x + ay * ay * w.y * w.y;
    return (sqrtf(1.0f + a2 * tan2) - 1.0f) / 2.0f;
}
RT_HD float mf_G1(V3 w, float ax, float ay) { return 1.0f / (1.0f + mf_lambda(w, ax, ay)); }
RT_HD float mf_G(V3 wo, V3 wi, float ax, float ay) { return 1.0f / (1.0f + mf_lambda(wo, ax, ay) + mf_lambda(wi, ax, ay)); }
RT_HD float mf_visible_distribution(V3 w, V3 wm, float ax, float ay) {
    float cos_theta = fabsf(w.z);
    return (mf_G1(w, ax, ay) / cos_theta) * mf_distribution(wm, ax, ay) * fabsf(dot(w, wm));
}
RT_HD_CALL V3 mf_sample_wm(V3 w, float ax, float ay, V2 u) {
    V3 wh = unit(mk3(ax * w.x, ay * w.y, w.z));
    if (wh.z < 0.0f) wh = -wh;
    V2 p = sample_unit_disk(u);
    V3 t1 = wh.z < 0.9999f ? cross(mk3(0, 0, 1), wh) : mk3(1, 0, 0);
    V3 t2 = cross(wh, t1);
    float h = sqrtf(1.0f - p.x * p.x);
    float offset = 0.5f * h * (1.0f - wh.z);
    float scale = 0.5f * (1.0f + wh.z);
    p = mk2(p.x, offset + scale * p.y);
    float pz = sqrtf(fmaxf(0.0f, 1.0f - sqmag(p)));
    V3 nh = p.x * t1 + p.y * t2 + pz * wh;
    return unit(mk3(ax * nh.x, ay * nh.y, fmaxf(1.0e-6f, nh.z)));
}

struct Bsdf {  // one interface (materials.rs:43-80 without the Layered arm)
    uint32_t kind;
    V3 albedo;      // Diffuse
    float eta;      // dielectrics
    V3 eta3, kappa; // conductors
    float ax, ay;
};
struct Layered {  // materials.rs:66-79
    Bsdf top, bottom;
    V3 albedo;
    float thickness, g;
    uint32_t n_samples, max_depth;
};

RT_HD bool bsdf_is_delta(const Bsdf& b) { return b.kind == B_SMOOTH_DIELECTRIC || b.kind == B_SMOOTH_CONDUCTOR; }
RT_HD uint32_t bsdf_components(const Bsdf& b) {  // materials.rs:682-691
    switch (b.kind) {
        case B_DIFFUSE: return NONSPEC_REFL;
        case B_SMOOTH_DIELECTRIC: return SPEC_REFL | SPEC_TRANS;
        case B_SMOOTH_CONDUCTOR: return SPEC_REFL;
        case B_ROUGH_CONDUCTOR: return NONSPEC_REFL;
        case B_ROUGH_DIELECTRIC: return NONSPEC_REFL | NONSPEC_TRANS;
        default: return 0;
    }
}

RT_HD float refl_pdf(V3 wo, V3 wi, float ax, float ay) {
    if (is_zero(wo + wi)) return 0.0f;
    V3 wm = unit(wo + wi);
    if (wm.z < 0.0f) wm = -wm;
    return mf_visible_distribution(wo, wm, ax, ay) / (4.0f * fabsf(dot(wo, wm)));
}
RT_HD V3 refl_bsdf(V3 wo, V3 wi, V3 eta, V3 kappa, float ax, float ay) {
    if (is_zero(wo + wi)) return mk3(0.0f);
    V3 wm = unit(wo + wi);
    float cos_theta = dot(wm, wi);
    V3 fr = fresnel_complex3(fabsf(cos_theta), eta, kappa);
    return mf_distribution(wm, ax, ay) * fr * mf_G(wo, wi, ax, ay) / (4.0f * wo.z * wi.z);
}
RT_HD_CALL float ts_pdf(V3 wo, V3 wi, float eta, float ax, float ay, uint32_t component) {
    bool refl = wo.z * wi.z > 0.0f;
    float eta_wm = !refl ? (wo.z > 0.0f ? eta : 1.0f / eta) : 1.0f;
    V3 wm = unit(wi * eta_wm + wo);
    if (wm.z < 0.0f) wm = -wm;
    if (wi.z == 0.0f || wo.z == 0.0f || is_zero(wm)) return 0.0f;
    if (dot(wm, wi) * wi.z < 0.0f || dot(wm, wo) * wo.z < 0.0f) return 0.0f;
    float R = fresnel_dielectric(dot(wo, wm), eta), T = 1.0f - R;
    float pr = (component & NONSPEC_REFL) ? R : 0.0f;
    float pt = (component & NONSPEC_TRANS) ? T : 0.0f;
    float ptot = pr + pt;
    if (refl) return (pr / ptot) * mf_visible_distribution(wo, wm, ax, ay) / (4.0f * fabsf(dot(wo, wm)));
    float dd = dot(wi, wm) + dot(wo, wm) / eta_wm;
    float denom = dd * dd;
    float dwm_dwi = fabsf(dot(wi, wm)) / denom;
    return (pt / ptot) * mf_visible_distribution(wo, wm, ax, ay) * dwm_dwi;
}
RT_HD_CALL V3 ts_bsdf(V3 wo, V3 wi, float eta, float ax, float ay) {
    bool refl = wo.z * wi.z > 0.0f;
    float eta_wm = !refl ? (wo.z > 0.0f ? eta : 1.0f / eta) : 1.0f;
    V3 wm = unit(wi * eta_wm + wo);
    if (wm.z < 0.0f) wm = -wm;
    if (wi.z == 0.0f || wo.z == 0.0f || is_zero(wm)) return mk3(0.0f);
    if (dot(wm, wi) * wi.z < 0.0f || dot(wm, wo) * wo.z < 0.0f) return mk3(0.0f);
    float F = fresnel_dielectric(dot(wo, wm), eta);
    if (refl) return mk3(mf_distribution(wm, ax, ay) * F * mf_G(wo, wi, ax, ay) / fabsf(4.0f * wo.z * wi.z));
    float dd = dot(wi, wm) + dot(wo, wm) / eta_wm;
    float denom = wi.z * wo.z * (dd * dd);
    return mk3(mf_distribution(wm, ax, ay) * (1.0f - F) * mf_G(wo, wi, ax, ay) * fabsf(dot(wi, wm) * dot(wo, wm) / denom) / (eta_wm * eta_wm));
}

RT_HD_CALL V3 bsdf_eval(const Bsdf& b, V3 wo, V3 wi) {  // materials.rs:125-169
    switch (b.kind) {
        case B_DIFFUSE: return wo.z * wi.z < 0.0f ? mk3(0.0f) : b.albedo / PI;
        case B_ROUGH_CONDUCTOR: return refl_bsdf(wo, wi, b.eta3, b.kappa, b.ax, b.ay);
        case B_ROUGH_DIELECTRIC: return ts_bsdf(wo, wi, b.eta, b.ax, b.ay);
        default: return mk3(0.0f);
    }
}
RT_HD_CALL float bsdf_pdf(const Bsdf& b, V3 wo, V3 wi, uint32_t component) {  // materials.rs:338-370
    switch (b.kind) {
        case B_DIFFUSE:
            if (!(component & NONSPEC_REFL)) return 0.0f;
            return wo.z * wi.z > 0.0f ? 1.0f / (2.0f * PI) : 0.0f;
        case B_ROUGH_CONDUCTOR:
            if (!(component & NONSPEC_REFL)) return 0.0f;
            return refl_pdf(wo, wi, b.ax, b.ay);
        case B_ROUGH_DIELECTRIC: return ts_pdf(wo, wi, b.eta, b.ax, b.ay, component);
        default: return 0.0f;
    }
}

RT_HD_CALL int bsdf_sample(const Bsdf& b, V3 wo, uint32_t component, Sampler& s, BsdfSample& out) {  // materials.rs:377-538
    switch (b.kind) {
        case B_DIFFUSE: {
            if (!(component & NONSPEC_REFL)) return S_INVALID;
            V3 wi = sample_cosine_hemisphere(s.uniform2());
            out.wi = wi; out.f = b.albedo / PI; out.pdf = wi.z / PI; out.component = NONSPEC_REFL;
            return validate_sample(out);
        }
        case B_SMOOTH_DIELECTRIC: {
            if (!(component & (SPEC_REFL | SPEC_TRANS))) return S_INVALID;
            V3 normal = mk3(0, 0, 1);
            float R = fresnel_dielectric(wo.z, b.eta), T = 1.0f - R;
            float pr = (component & SPEC_REFL) ? R : 0.0f, pt = (component & SPEC_TRANS) ? T : 0.0f;
            float ptot = pr + pt;
            float smp = s.uniform();
            if (smp * ptot < pr) {
                V3 rd = reflect(wo, normal);
                out.wi = rd; out.f = mk3(R / fabsf(rd.z)); out.pdf = R / ptot; out.component = SPEC_REFL;
            } else {
                V3 rd;
                if (!refract(b.eta, wo, normal, rd)) return S_INVALID;
                float e = wo.z < 0.0f ? 1.0f / b.eta : b.eta;
                out.wi = rd; out.f = mk3((T / fabsf(rd.z)) / (e * e)); out.pdf = T / ptot; out.component = SPEC_TRANS;
            }
            return validate_sample(out);
        }
        case B_SMOOTH_CONDUCTOR: {
            V3 rd = reflect(wo, mk3(0, 0, 1));
            V3 fr = fresnel_complex3(wo.z, b.eta3, b.kappa);
            out.wi = rd; out.f = mk3(fr.x / wo.z, fr.y / wo.z, fr.z / wo.z); out.pdf = 1.0f; out.component = SPEC_REFL;
            return validate_sample(out);
        }
        case B_ROUGH_CONDUCTOR: {
            V3 wm = mf_sample_wm(wo, b.ax, b.ay, s.uniform2());
            V3 wi = reflect(wo, wm);
            if (wo.z * wi.z < 0.0f) return S_NULL;
            out.wi = wi; out.pdf = refl_pdf(wo, wi, b.ax, b.ay); out.f = refl_bsdf(wo, wi, b.eta3, b.kappa, b.ax, b.ay);
            out.component = NONSPEC_REFL;
            return validate_sample(out);
        }
        default: {  // B_ROUGH_DIELECTRIC
            V3 wm = mf_sample_wm(wo, b.ax, b.ay, s.uniform2());
            float R = fresnel_dielectric(dot(wo, wm), b.eta), T = 1.0f - R;
            float pr = (component & REFLECTION) == REFLECTION ? R : 0.0f;
            float pt = (component & TRANSMISSION) == TRANSMISSION ? T : 0.0f;
            float ptot = pr + pt;
            V3 wi;
            bool reflected;
            if (s.uniform() * ptot < pr) {
                wi = reflect(wo, wm);
                if (wo.z * wi.z < 0.0f) return S_NULL;
                reflected = true;
            } else {
                if (!refract(b.eta, wo, wm, wi)) return S_INVALID;
                if (wo.z * wi.z > 0.0f || wi.z == 0.0f) return S_NULL;
                reflected = false;
            }
            out.wi = wi; out.pdf = ts_pdf(wo, wi, b.eta, b.ax, b.ay, component); out.f = ts_bsdf(wo, wi, b.eta, b.ax, b.ay);
            out.component = reflected ? NONSPEC_REFL : NONSPEC_TRANS;
            return validate_sample(out);
        }
    }
}

// ---- Henyey-Greenstein, materials.rs:1477-1535 -----------------------------------------------------
RT_HD float hg_phase(float cos_theta, float g) {
    float denom = 1.0f + g * g + 2.0f * g * cos_theta;
    return FRAC_1_PI * 0.25f * (1.0f - g * g) / (denom * sqrtf(denom));
}
struct PhaseSample { V3 wi; float p, pdf; };
RT_HD PhaseSample phase_sample(V3 wo, float g, V2 u) {
    float cos_theta;
    if (fabsf(g) < 1.0e-3f) cos_theta = 1.0f - 2.0f * u.x;
    else {
        float term = (1.0f - g * g) / (1.0f + g - 2.0f * g * u.x);
        cos_theta = -1.0f / (2.0f * g) * (1.0f + g * g - term * term);
    }
    float phi = 2.0f * PI * u.y;
    float sin_theta = sqrtf(1.0f - cos_theta * cos_theta);
    V3 d = mk3(cosf(phi) * sin_theta, sinf(phi) * sin_theta, cos_theta);
    V3 wx, wy;
    make_orthonormal_basis(wo, wx, wy);
    PhaseSample ps;
    ps.wi = d.x * wx + d.y * wy + d.z * wo;
    ps.p = ps.pdf = hg_phase(cos_theta, g);
    return ps;
}
RT_HD float phase_p(V3 wo, V3 wi, float g) { return hg_phase(dot(wo, wi), g); }

RT_HD float layer_tr(float dz, V3 w) { return expf(-fabsf(dz / w.z)); }  // materials.rs:84-87

// LayeredBsdf::sample_bsdf, materials.rs:540-666
RT_HD_CALL int layered_sample(const Layered& L, V3 wo, Sampler& s, BsdfSample& out) {
    bool flip_wi = false;
    if (wo.z < 0.0f) { wo = -wo; flip_wi = true; }
    BsdfSample enter;
    int st = bsdf_sample(L.top, wo, ALL_COMPONENTS, s, enter);
    if (st != S_VALID) return st;
    if (enter.component & REFLECTION) {
        out = enter;
        if (flip_wi) out.wi = -enter.wi;
        return validate_sample(out);
    }
    bool specular_path = (enter.component & SPECULAR) != 0;
    V3 w = enter.wi;
    V3 f = enter.f * fabsf(enter.wi.z);
    float pdf = enter.pdf;
    float z = L.thickness;
    for (uint32_t depth = 0; depth < L.max_depth; depth++) {
        float rr_beta = max_component(f) / pdf;
        if (depth > 3 && rr_beta < 0.25f) {
            float q = fmaxf(0.0f, 1.0f - rr_beta);
            if (s.uniform() < q) return S_NULL;
            pdf *= 1.0f - q;
        }
        if (w.z == 0.0f) return S_NULL;
        if (!is_zero(L.albedo)) {
            float dz = sample_exponential(s.uniform(), 1.0f / fabsf(w.z));
            float zp = w.z > 0.0f ? z + dz : z - dz;
            if (zp == z) return S_INVALID;
            if (0.0f < zp && zp < L.thickness) {
                PhaseSample ps = phase_sample(-w, L.g, s.uniform2());
                if (ps.wi.z == 0.0f) return S_NULL;
                f *= L.albedo * ps.p;
                pdf *= ps.pdf;
                specular_path = false;
                w = ps.wi;
                z = zp;
                continue;
            }
            z = rs_clamp(zp, 0.0f, L.thickness);
        } else {
            z = (z == L.thickness) ? 0.0f : L.thickness;
            f *= layer_tr(L.thickness, w);
        }
        const Bsdf& iface = (z == 0.0f) ? L.bottom : L.top;
        BsdfSample is;
        st = bsdf_sample(iface, -w, ALL_COMPONENTS, s, is);
        if (st != S_VALID) return st;
        f *= is.f;
        pdf *= is.pdf;
        specular_path = specular_path && (is.component & SPECULAR) != 0;
        w = is.wi;
        if (is.component & TRANSMISSION) {
            bool same_dir = wo.z * w.z > 0.0f;
            uint32_t comp = same_dir ? (specular_path ? SPEC_REFL : NONSPEC_REFL) : (specular_path ? SPEC_TRANS : NONSPEC_TRANS);
            if (flip_wi) w = -w;
            out.wi = w; out.f = f; out.pdf = pdf; out.component = comp;
            return validate_sample(out);
        }
        f *= fabsf(is.wi.z);
    }
    return S_NULL;
}

// LayeredBsdf::evaluate_bsdf, materials.rs:171-333. The random walk draws from a private sampler seeded by
// FxHash(wi bits, wo bits) (materials.rs:209-214) and never touches the path stream.
RT_HD_CALL V3 layered_eval(const Layered& L, V3 wo, V3 wi) {
    V3 f = mk3(0.0f);
    if (wo.z < 0.0f) { wo = -wo; wi = -wi; }
    const Bsdf& enter_if = L.top;
    const bool exit_bottom = wi.z < 0.0f;
    const Bsdf& exit_if = exit_bottom ? L.bottom : L.top;
    const Bsdf& non_exit_if = exit_bottom ? L.top : L.bottom;
    const float exit_z = exit_bottom ? 0.0f : L.thickness;
    if (!(bsdf_components(enter_if) & TRANSMISSION) || !(bsdf_components(exit_if) & TRANSMISSION)) return mk3(0.0f);
    if (wo.z * wi.z > 0.0f) f += (float)L.n_samples * bsdf_eval(enter_if, wo, wi);
    FxHasher h;
    h.write_u32(f2u(wi.x)); h.write_u32(f2u(wi.y)); h.write_u32(f2u(wi.z));
    h.write_u32(f2u(wo.x)); h.write_u32(f2u(wo.y)); h.write_u32(f2u(wo.z));
    Sampler s;
    s.init_one_off(h.finish());
    const bool exit_delta = bsdf_is_delta(exit_if), non_exit_delta = bsdf_is_delta(non_exit_if);
    for (uint32_t i = 0; i < L.n_samples; i++) {
        BsdfSample enter, exitS;
        if (bsdf_sample(enter_if, wo, TRANSMISSION, s, enter) != S_VALID) continue;
        if (bsdf_sample(exit_if, wi, TRANSMISSION, s, exitS) != S_VALID) continue;
        V3 beta = exitS.f * fabsf(exitS.wi.z) / exitS.pdf;
        float z = L.thickness;
        V3 w = enter.wi;
        for (uint32_t depth = 0; depth < L.max_depth; depth++) {
            if (depth > 3 && max_component(beta) < 0.25f) {
                float q = fmaxf(0.0f, max_component(beta));
                if (s.uniform() < q) break;
                beta /= 1.0f - q;
            }
            if (is_zero(L.albedo)) {
                z = (z == L.thickness) ? 0.0f : L.thickness;
                beta *= layer_tr(L.thickness, w);
            } else {
                float dz = sample_exponential(s.uniform(), 1.0f / fabsf(w.z));
                float zp = w.z > 0.0f ? z + dz : z - dz;
                if (0.0f < zp && zp < L.thickness) {
                    float ph = phase_p(-w, -exitS.wi, L.g);
                    float wt = exit_delta ? 1.0f : power_heuristic(1, exitS.pdf, 1, ph);
                    f += beta * L.albedo * ph * wt * layer_tr(zp - exit_z, exitS.wi) * exitS.f / exitS.pdf;
                    PhaseSample ps = phase_sample(-w, L.g, s.uniform2());
                    beta *= L.albedo * ps.p / ps.pdf;
                    w = ps.wi;
                    z = zp;
                    bool facing_exit = (z < exit_z && w.z > 0.0f) || (z > exit_z && w.z < 0.0f);
                    if (!exit_delta && facing_exit) {
                        V3 exit_f = bsdf_eval(exit_if, -w, wi);
                        if (!is_zero(exit_f)) {
                            float exit_pdf = bsdf_pdf(exit_if, -w, wi, TRANSMISSION);
                            float wt2 = power_heuristic(1, ps.pdf, 1, exit_pdf);
                            f += beta * layer_tr(zp - exit_z, ps.wi) * exit_f * wt2;
                        }
                    }
                    continue;
                }
                z = rs_clamp(zp, 0.0f, L.thickness);
            }
            if (z == exit_z) {
                BsdfSample rs;
                if (bsdf_sample(exit_if, -w, REFLECTION, s, rs) != S_VALID) break;
                beta *= rs.f * fabsf(rs.wi.z) / rs.pdf;
                w = rs.wi;
            } else {
                if (!non_exit_delta) {
                    float wt = power_heuristic(1, exitS.pdf, 1, bsdf_pdf(non_exit_if, -w, -exitS.wi, REFLECTION));
                    f += beta * bsdf_eval(non_exit_if, -w, -exitS.wi) * fabsf(exitS.wi.z) * wt * layer_tr(L.thickness, exitS.wi) * exitS.f / exitS.pdf;
                }
                BsdfSample ns;
                if (bsdf_sample(non_exit_if, -w, REFLECTION, s, ns) != S_VALID) break;
                beta *= ns.f * fabsf(ns.wi.z) / ns.pdf;
                w = ns.wi;
                if (!exit_delta) {
                    V3 exit_f = bsdf_eval(exit_if, -w, wi);
                    if (!is_zero(exit_f)) {
                        float exit_pdf = bsdf_pdf(exit_if, -w, wi, ALL_COMPONENTS);
                        float wt = non_exit_delta ? 1.0f : power_heuristic(1, ns.pdf, 1, exit_pdf);
                        f += beta * layer_tr(L.thickness, ns.wi) * exit_f * wt;
                    }
                }
            }
        }
    }
    return f / (float)L.n_samples;
}

// The material at a hit: one interface, or a layered stack.
struct Surface {
    bool layered;
    Bsdf b;
    Layered l;
};
RT_HD bool surface_is_delta(const Surface& s) { return !s.layered && bsdf_is_delta(s.b); }
// Diffuse (the only material the glTF importer produces, scene.rs:406-407) is evaluated inline; everything
// else goes through the out-of-line general BSDF code.
RT_HD V3 surface_eval(const Surface& s, V3 wo, V3 wi) {
    if (!s.layered && s.b.kind == B_DIFFUSE) return wo.z * wi.z < 0.0f ? mk3(0.0f) : s.b.albedo / PI;
    return s.layered ? layered_eval(s.l, wo, wi) : bsdf_eval(s.b, wo, wi);
}
RT_HD int surface_sample(const Surface& s, V3 wo, Sampler& smp, BsdfSample& out) {
    if (!s.layered && s.b.kind == B_DIFFUSE) {
        V3 wi = sample_cosine_hemisphere(smp.uniform2());
        out.wi = wi; out.f = s.b.albedo / PI; out.pdf = wi.z / PI; out.component = NONSPEC_REFL;
        return validate_sample(out);
    }
    return s.layered ? layered_sample(s.l, wo, smp, out) : bsdf_sample(s.b, wo, ALL_COMPONENTS, smp, out);
}

constexpr float MINIMUM_ROUGHNESS = 1.0e-3f;  // materials.rs:1538-1541

RT_HD void zero_bsdf(Bsdf& b) {
    b.kind = B_DIFFUSE; b.albedo = mk3(0.0f); b.eta = 1.0f; b.eta3 = mk3(0.0f); b.kappa = mk3(0.0f); b.ax = b.ay = 0.0f;
}

// CpuMaterial::get_bsdf, materials.rs:823-955
RT_HD void get_surface(const SceneD& sc, const MaterialD& m, const MatCtx& c, Surface& out) {
    out.layered = false;
    Bsdf& b = out.b;
    zero_bsdf(b);
    switch (m.kind) {
        case 0: b.kind = B_DIFFUSE; b.albedo = xyz(tex(sc, m.albedo, c)); break;
        case 1: b.kind = B_SMOOTH_DIELECTRIC; b.eta = tex(sc, m.eta, c).x; break;
        case 2: b.kind = B_SMOOTH_CONDUCTOR; b.eta3 = xyz(tex(sc, m.eta, c)); b.kappa = xyz(tex(sc, m.kappa, c)); break;
        case 4: {
            b.eta3 = xyz(tex(sc, m.eta, c)); b.kappa = xyz(tex(sc, m.kappa, c));
            V4 r = tex(sc, m.roughness, c);
            b.ax = m.remap_roughness ? sqrtf(r.x) : r.x;
            b.ay = m.remap_roughness ? sqrtf(r.y) : r.y;
            b.kind = fmaxf(b.ax, b.ay) < MINIMUM_ROUGHNESS ? B_SMOOTH_CONDUCTOR : B_ROUGH_CONDUCTOR;
            break;
        }
        case 3: {
            b.eta = tex(sc, m.eta, c).x;
            V4 r = tex(sc, m.roughness, c);
            b.ax = m.remap_roughness ? sqrtf(r.x) : r.x;
            b.ay = m.remap_roughness ? sqrtf(r.y) : r.y;
            b.kind = fmaxf(b.ax, b.ay) < MINIMUM_ROUGHNESS ? B_SMOOTH_DIELECTRIC : B_ROUGH_DIELECTRIC;
            break;
        }
        default: {  // CoatedDiffuse
            out.layered = true;
            Layered& l = out.l;
            zero_bsdf(l.bottom);
            l.bottom.kind = B_DIFFUSE;
            l.bottom.albedo = xyz(tex(sc, m.albedo, c));
            zero_bsdf(l.top);
            l.top.eta = tex(sc, m.eta, c).x;
            l.top.kind = B_SMOOTH_DIELECTRIC;
            if (m.roughness != NONE) {
                V4 r = tex(sc, m.roughness, c);
                l.top.ax = m.remap_roughness ? sqrtf(r.x) : r.x;
                l.top.ay = m.remap_roughness ? sqrtf(r.y) : r.y;
                if (!(fmaxf(l.top.ax, l.top.ay) < MINIMUM_ROUGHNESS)) l.top.kind = B_ROUGH_DIELECTRIC;
            }
            l.n_samples = 8;
            l.max_depth = 8;
            l.thickness = tex(sc, m.thickness, c).x;
            l.albedo = xyz(tex(sc, m.coat_albedo, c));
            l.g = 0.0f;
        }
    }
}
// Scenes whose materials are all Diffuse (everything the glTF importer produces, scene.rs:406-407) are shaded by a
// kernel instantiated on this 3-float surface instead of the general Surface (two Bsdf records + the layered stack):
// same arithmetic as the Diffuse branches above, a third fewer live registers in k_shade.
struct DiffuseSurface { V3 albedo; };
RT_HD bool surface_is_delta(const DiffuseSurface&) { return false; }
RT_HD V3 surface_eval(const DiffuseSurface& s, V3 wo, V3 wi) { return wo.z * wi.z < 0.0f ? mk3(0.0f) : s.albedo / PI; }
RT_HD int surface_sample(const DiffuseSurface& s, V3, Sampler& smp, BsdfSample& out) {
    V3 wi = sample_cosine_hemisphere(smp.uniform2());
    out.wi = wi; out.f = s.albedo / PI; out.pdf = wi.z / PI; out.component = NONSPEC_REFL;
    return validate_sample(out);
}
RT_HD void get_surface(const SceneD& sc, const MaterialD& m, const MatCtx& c, DiffuseSurface& out) { out.albedo = xyz(tex(sc, m.albedo, c)); }
// Diffuse material by index: a constant albedo comes from the per-material table (one load instead of the dependent
// material -> texture -> value chain); the value is the one tex() returns for a constant texture
RT_HD void get_surface_of(const SceneD& sc, uint32_t material, const MatCtx& c, DiffuseSurface& out) {
    if (sc.mat_const) {
        const float4 mc = ldg(sc.mat_const + material);
        if (mc.w == 1.0f) { out.albedo = xyz(mc); return; }
    }
    get_surface(sc, sc.materials[material], c, out);
}
RT_HD void get_surface_of(const SceneD& sc, uint32_t material, const MatCtx& c, Surface& out) { get_surface(sc, sc.materials[material], c, out); }

RT_HD bool get_mip_level(const SceneD& sc, const MaterialD& m, const MatCtx& c, float& level) {  // materials.rs:957-968
    if (m.kind != 0) return false;
    return texture_mip_level(sc, m.albedo, c, level);
}
RT_HD_CALL V3 get_albedo(const SceneD& sc, const MaterialD& m, const MatCtx& c) {  // materials.rs:970-987
    if (m.kind == 0 || m.kind == 5) return xyz(tex(sc, m.albedo, c));
    return mk3(1.0f);
}

}  // namespace rt
