// ref_units.cpp — C entry points over the reference's OWN C++ restatement of the CPU renderer's math
// (/root/reference/crates/raytracing-optix/csrc/kernels/*.hpp, functions annotated `@raytracing_cpu::...`), compiled
// unmodified for the host (oracle/ref_shim/preamble.h). TEST INFRASTRUCTURE: built into oracle/_ref/libref_units.so by
// oracle/ref_shim/Makefile; tests/test_oracle_ref.py checks oracle/oracle.cpp against it on random inputs. Never part of
// the product. The unit kinds and row layouts are shared with oracle_unit_batch (oracle/oracle.cpp).
#include "materials.hpp"   // pulls kernel_math, kernel_types, sample, texture, pathtracer
#include "geometry.hpp"
#include "camera.hpp"

extern "C" {
PathtracerPipelineParams pipeline_params;   // the __constant__ block texture.hpp / lights.hpp refer to
}

namespace {
constexpr int IN_W = 24, OUT_W = 8;
float3 f3(const float* p) { return make_float3(p[0], p[1], p[2]); }
void put3(float* o, float3 v) { o[0] = v.x; o[1] = v.y; o[2] = v.z; }
}

extern "C" __attribute__((visibility("default"))) int ref_unit_io_width(int which) { return which == 0 ? IN_W : OUT_W; }

extern "C" __attribute__((visibility("default"))) int ref_unit_batch(int kind, unsigned int n, const float* in, float* out) {
    using namespace materials;
    for (unsigned int i = 0; i < n; i++) {
        const float* a = in + (size_t)i * IN_W;
        float* o = out + (size_t)i * OUT_W;
        for (int k = 0; k < OUT_W; k++) o[k] = 0.0f;
        switch (kind) {
            case 0: o[0] = fresnel_dielectric(a[0], a[1]); break;
            case 1: o[0] = fresnel_complex(a[0], complex(a[1], a[2])); break;
            case 2: { auto r = refract(a[0], f3(a + 1), f3(a + 4)); o[0] = r ? 1.0f : 0.0f; if (r) put3(o + 1, *r); break; }
            case 3: put3(o, reflect(f3(a), f3(a + 3))); break;
            case 4: o[0] = microfacet::distribution(f3(a), a[3], a[4]); break;
            case 5: o[0] = microfacet::lambda(f3(a), a[3], a[4]); break;
            case 6: o[0] = microfacet::G1(f3(a), a[3], a[4]); break;
            case 7: o[0] = microfacet::G(f3(a), f3(a + 3), a[6], a[7]); break;
            case 8: o[0] = microfacet::visible_distribution(f3(a), f3(a + 3), a[6], a[7]); break;
            case 9: put3(o, microfacet::sample_wm(f3(a), a[3], a[4], make_float2(a[5], a[6]))); break;
            case 10: { OptixBsdfRoughConductor b{f3(a), f3(a + 3), a[6], a[7]}; put3(o, evaluate_bsdf(b, f3(a + 8), f3(a + 11))); break; }
            case 11: { OptixBsdfRoughConductor b{f3(a), f3(a + 3), a[6], a[7]}; o[0] = evaluate_pdf(b, f3(a + 8), f3(a + 11), BsdfComponentFlags::ALL()); break; }
            case 12: { OptixBsdfRoughDielectric b{a[0], a[1], a[2]}; put3(o, evaluate_bsdf(b, f3(a + 3), f3(a + 6))); break; }
            case 13: { OptixBsdfRoughDielectric b{a[0], a[1], a[2]}; o[0] = evaluate_pdf(b, f3(a + 3), f3(a + 6), BsdfComponentFlags::ALL()); break; }
            case 14: { OptixBsdfDiffuse b{f3(a)}; put3(o, evaluate_bsdf(b, f3(a + 3), f3(a + 6))); break; }
            case 15: { auto xy = geometry::make_orthonormal_basis(f3(a)); put3(o, xy.first); put3(o + 3, xy.second); break; }
            case 16: o[0] = geometry::tri_area(f3(a), f3(a + 3), f3(a + 6)); break;
            case 17: { float2 d = sample::sample_unit_disk(make_float2(a[0], a[1])); o[0] = d.x; o[1] = d.y; break; }
            case 18: { float2 d = sample::sample_unit_disk_concentric(make_float2(a[0], a[1])); o[0] = d.x; o[1] = d.y; break; }
            case 19: put3(o, sample::sample_cosine_hemisphere(make_float2(a[0], a[1]))); break;
            case 20: o[0] = sample::sample_exponential(a[0], a[1]); break;
            case 21: { complex c = complex(a[0], a[1]).sqrt(); o[0] = c.real; o[1] = c.imag; break; }
            case 22: {   // SmoothConductor::sample_bsdf draws nothing from the sampler
                OptixBsdfSmoothConductor b{f3(a), f3(a + 3)};
                sample::OptixSampler s = sample::OptixSampler::one_off_sampler(1);
                BsdfSample bs = sample_bsdf(b, f3(a + 6), BsdfComponentFlags::ALL(), s);
                put3(o, bs.wi); put3(o + 3, bs.bsdf); o[6] = bs.pdf; o[7] = bs.valid ? 1.0f : 0.0f;
                break;
            }
            case 23: { Matrix4x4 m; for (int k = 0; k < 16; k++) m.m[k] = a[k]; put3(o, matrix4x4_apply_point(m, f3(a + 16))); break; }
            case 24: { Matrix4x4 m; for (int k = 0; k < 16; k++) m.m[k] = a[k]; put3(o, matrix4x4_apply_vector(m, f3(a + 16))); break; }
            default: return 1;
        }
    }
    return 0;
}

// camera.hpp generate_ray without a sampler (pixel corner + (x, y) as given): kind 0 orthographic, 1 pinhole.
// out: origin.xyz, direction.xyz per ray.
extern "C" __attribute__((visibility("default"))) int ref_camera_rays(int kind, const float* raster_to_camera, const float* camera_to_world,
                                                                       unsigned int n, const unsigned int* xy, float* out) {
    Camera cam{};
    cam.camera_type.kind = kind == 0 ? Orthographic : PinholePerspective;
    for (int k = 0; k < 16; k++) { cam.raster_to_camera.forward.m[k] = raster_to_camera[k]; cam.camera_to_world.forward.m[k] = camera_to_world[k]; }
    for (unsigned int i = 0; i < n; i++) {
        Ray r = generate_ray(cam, xy[2 * i], xy[2 * i + 1], nullptr);
        put3(out + 6 * (size_t)i, r.origin);
        put3(out + 6 * (size_t)i + 3, r.direction);
    }
    return 0;
}
