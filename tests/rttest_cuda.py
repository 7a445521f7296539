"""`rttest`-style visual regression of the cuda backend against CPU-blessed references — TEST INFRASTRUCTURE (uses the
oracle). Mirrors visual-testing/src/rttest: for every test of the suite (tests/tests.toml of the reference = the
builtin scenes; external assets bathroom / sponza are absent) the scene is rendered through the driver
(raytracing_cuda.cli, `--backend cuda`) into an EXR, the reference EXR is blessed from the CPU restatement of the
reference renderer with the same arguments, and the two are compared with rttest's own metric (MSE / max diff over
all channels, diff.py:64-89) plus FLIP and the mean-luminance z-score the north star names. Because the reference
ships no blessed images, the pass criterion is relative: the cuda frame must differ from the CPU frame (same seed) by
no more than two CPU frames rendered with different seeds differ from each other (the noise floor at equal spp).

    python tests/rttest_cuda.py --out profiles/r1_visual_regression.md      # on a GPU box
"""
import argparse
import os
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)

SUITE = ["sphere", "cube", "cube_orthographic", "checkered_plane", "dielectric", "metal", "rough_metal", "rough_dielectric",
         "out_of_focus_sphere"]
GLTF_SUITE = [("cb", 256, 256), ("cbbunny_area_light_transforms", 320, 180), ("cb_texture", 320, 180)]


def _bless(rc, oracle, scene, settings, path, seed=None):
    import copy
    st = copy.copy(settings)
    if seed is not None:
        st.seed = seed
    out, _ = oracle.render(scene, st, num_threads=os.cpu_count() or 1)
    rc.exr.write_exr(path, rc.exr.channels_of_render_output(out, rc.AovFlags(st.outputs)))
    return out


def run_suite(rc, oracle, renderer_args=("-s", "4", "-l", "1"), workdir=None, gltf=True):
    from raytracing_cuda import cli, imagecmp
    workdir = workdir or tempfile.mkdtemp(prefix="rttest_cuda_")
    results = []
    cases = [(n, None) for n in SUITE] + ([(n, (w, h)) for n, w, h in GLTF_SUITE] if gltf else [])
    for name, size in cases:
        out_path = os.path.join(workdir, "output", f"{name}.exr")
        ref_path = os.path.join(workdir, "reference", f"{name}.exr")
        for d in (os.path.dirname(out_path), os.path.dirname(ref_path)):
            os.makedirs(d, exist_ok=True)
        flags = cli.build_parser().parse_args(["--scene-name", "sphere"] + list(renderer_args))
        if size is None:   # builtin scene: through the driver, exactly like `rttest` invokes `dist/cli`
            t = [t for t in rc.test_scenes.all_test_scenes() if t.name == name][0]
            scene, settings = t.scene_func(), t.settings_func()
            argv = ["--scene-name", name, "--output-dir", os.path.dirname(out_path), "-o", f"{name}.exr"] + list(renderer_args) + ["full"]
            assert cli.main(argv) == 0, argv
        else:              # glTF fixture (.npz of the imported scene; the .glb files live in the reference checkout only)
            from conftest import load_scene
            scene, settings = load_scene(name, *size), rc.RaytracerSettings()
        for field, value in (("max_ray_depth", flags.ray_depth), ("light_sample_count", flags.light_samples), ("samples_per_pixel", flags.spp)):
            if value is not None:
                setattr(settings, field, value)
        if size is not None:
            out = rc.render(scene, settings)
            rc.exr.write_exr(out_path, rc.exr.channels_of_render_output(out, rc.AovFlags(settings.outputs)))
        _bless(rc, oracle, scene, settings, ref_path)
        got, w, h = rc.exr.read_exr(out_path)
        ref, rw, rh = rc.exr.read_exr(ref_path)
        assert (w, h) == (rw, rh) and sorted(got) == sorted(ref), f"{name}: resolution / channel mismatch"
        names = sorted(got)
        g = np.stack([got[c] for c in names], axis=-1)
        r = np.stack([ref[c] for c in names], axis=-1)
        mse, maxd = imagecmp.mse_maxdiff(np.nan_to_num(g), np.nan_to_num(r))
        row = {"scene": name, "channels": names, "mse": mse, "max_diff": maxd}
        if all(c in got for c in "RGB"):
            gb = np.stack([got[c] for c in "RGB"], axis=-1)
            rb = np.stack([ref[c] for c in "RGB"], axis=-1)
            other = _bless(rc, oracle, scene, settings, os.path.join(workdir, "reference", f"{name}_seed7.exr"), seed=7).beauty
            exposure = 1.0 / max(float(np.percentile(imagecmp.luminance(rb), 99)), 1e-6)
            row["flip_vs_cpu"] = imagecmp.flip(rb, gb, exposure)
            row["flip_noise_floor"] = imagecmp.flip(rb, other, exposure)
            row["mse_noise_floor"] = imagecmp.mse_maxdiff(np.nan_to_num(rb), np.nan_to_num(other))[0]
            row["mse_beauty"] = imagecmp.mse_maxdiff(np.nan_to_num(gb), np.nan_to_num(rb))[0]
            row["luminance_z_vs_other_seed"] = imagecmp.mean_luminance_z(np.nan_to_num(gb), np.nan_to_num(other), settings.samples_per_pixel)
            row["passed"] = bool(row["flip_vs_cpu"] <= max(row["flip_noise_floor"], 1e-3) and row["mse_beauty"] <= max(row["mse_noise_floor"], 1e-10))
        else:
            row["passed"] = bool(maxd <= 1e-3 and mse <= 1e-8)   # deterministic AOV planes (a few silhouette pixels may flip)
        results.append(row)
    return results


def markdown(results, renderer_args):
    lines = ["| scene | channels | MSE (rttest metric) | max diff | FLIP cuda vs CPU | FLIP CPU vs CPU (other seed) | luminance z vs other seed | pass |",
             "|---|---|---|---|---|---|---|---|"]
    for r in results:
        f = lambda k: (f"{r[k]:.3g}" if k in r else "–")
        lines.append(f"| {r['scene']} | {','.join(r['channels'])} | {r['mse']:.3g} | {r['max_diff']:.3g} | {f('flip_vs_cpu')} | "
                     f"{f('flip_noise_floor')} | {f('luminance_z_vs_other_seed')} | {'yes' if r['passed'] else 'NO'} |")
    return f"renderer args: `{' '.join(renderer_args)}`\n\n" + "\n".join(lines) + "\n"


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--out")
    ap.add_argument("renderer_args", nargs="*", default=["-s", "16", "-l", "1"])
    a = ap.parse_args()
    import raytracing_cuda as rc
    import oracle_py
    oracle_py.build()
    res = run_suite(rc, oracle_py, a.renderer_args)
    md = markdown(res, a.renderer_args)
    print(md)
    if a.out:
        open(a.out, "w").write("# rttest-style visual regression: cuda backend vs CPU-blessed references\n\n" + md)
    sys.exit(0 if all(r["passed"] for r in res) else 1)
