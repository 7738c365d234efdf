"""Golden vectors FROM THE REFERENCE ITSELF (run in the build container only).

    python tests/golden/make_ref_pins.py

The reference's LBM tunnel is JavaScript + GLSL inside pages/airfoil_flow_lbm_aerolab.html; no
JavaScript engine or browser exists in the build image.  This script therefore executes the
reference's OWN SOURCE TEXT -- read from /root/reference at run time, never copied into this
repository -- with the purpose-built minimal interpreters in tests/refexec/ (JavaScript with
float64 semantics, GLSL with strict fp32 semantics) and commits only the numeric outputs:

  ref_pins.json / ref_pins.npz
    geometry   naca4/clarkY/SHAPES, rotate, panelise, rasterMask  (HTML:99-182)   -> panel nodes, masks
    init       equilibriumInitData                                 (HTML:474-490)  -> the 9+3 fp32 values
    step       STEP_FS_SRC.main run as a fragment shader per cell  (HTML:222-360)  -> populations + macro
    stats      updateFieldsFromMacro                               (HTML:596-614)  -> maxS, cpMin, cpMax
    forces     computeForces incl. EMAs and separation             (HTML:649-700)  -> CL/CD/sep series
    render     RENDER_FS_SRC.main, three field modes               (HTML:362-422)  -> RGBA8 images

tests/test_reference_pins.py checks the CPU oracle against these files (everywhere) and, when
/root/reference is present, re-executes a subset live.
"""
import hashlib
import json
import os
import sys
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, os.path.join(ROOT, "tests"))

from refexec import extract, glslrun, jsrun  # noqa: E402

f32 = np.float32


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def js_tunnel(page, nx, ny):
    """An interpreter holding the page's geometry / host functions for an nx x ny lattice."""
    I = jsrun.Interp()
    I.run(extract.js_statement(page, "const DX0="))
    I.globals.declare("NX", float(nx))
    I.globals.declare("NY", float(ny))
    for pre in ("const CHORD_L =", "const TAU =", "const NU_L =", "const NP=", "let U0=", "const FACE_DX=",
                "const macro=", "const Ufield=", "const Vfield=", "const CpField=", "let maxS=",
                "let CLsmooth="):
        I.run(extract.js_statement(page, pre))
    I.globals.declare("sol", None)
    for fn in ("naca4", "clarkY", "rotate", "panelise", "rasterMask", "equilibriumInitData",
               "updateFieldsFromMacro", "computeForces"):
        I.run(extract.js_function(page, fn))
    I.run(extract.js_statement(page, "const SHAPES="))
    return I


def geometry(page, cases):
    out, arrays = [], {}
    for k, (shape, alpha, nx, ny) in enumerate(cases):
        t0 = time.time()
        I = js_tunnel(page, nx, ny)
        base = I.globals.get("SHAPES")[shape]()
        pan = I.call("panelise", I.call("rotate", base, float(alpha)))
        mask = I.call("rasterMask", pan["xp"], pan["yp"]).a.reshape(ny, nx)
        arrays[f"geom{k}_xp"] = np.array(pan["xp"])
        arrays[f"geom{k}_yp"] = np.array(pan["yp"])
        arrays[f"geom{k}_base"] = np.array(base)
        out.append(dict(shape=shape, alpha=alpha, nx=nx, ny=ny, solid=int((mask > 0).sum()), sha256=sha(mask)))
        print("geometry", out[-1], f"{time.time() - t0:.1f}s", flush=True)
    return out, arrays


def user_coords_geometry(page, coords, alpha, nx, ny):
    I = js_tunnel(page, nx, ny)
    pan = I.call("panelise", I.call("rotate", [list(map(float, p)) for p in coords], float(alpha)))
    mask = I.call("rasterMask", pan["xp"], pan["yp"]).a.reshape(ny, nx)
    return dict(alpha=alpha, nx=nx, ny=ny, solid=int((mask > 0).sum()), sha256=sha(mask))


def run_tunnel(page, nx, ny, u0, tau, mask, frames, steps_per_frame, forces_every):
    """The page's frame loop: simStep x n, readMacro, updateFieldsFromMacro, computeForces."""
    I = js_tunnel(page, nx, ny)
    I.globals.set("U0", float(u0))
    init = I.call("equilibriumInitData", float(u0))
    tex = [init[k].a.reshape(ny, nx, 4).copy() for k in ("dA", "dB", "dC")]
    init_vals = np.concatenate([tex[0][0, 0], tex[1][0, 0], tex[2][0, 0]])
    step = glslrun.Shader(extract.shader(page, "STEP_FS_SRC"))
    from refexec.jsrun import TypedArray
    IN = TypedArray("Uint8Array", nx * ny)
    IN.a[:] = mask.reshape(-1)
    I.globals.set("sol", {"IN": IN})
    mask_tex = glslrun.Sampler((mask.astype(np.float32) / np.float32(255.0)).reshape(ny, nx, 1))
    series = []
    nsteps = 0
    for frame in range(1, frames + 1):
        for _ in range(steps_per_frame):
            t0 = time.time()
            uni = dict(texA=glslrun.Sampler(tex[0]), texB=glslrun.Sampler(tex[1]), texC=glslrun.Sampler(tex[2]),
                       texMask=mask_tex, texel=glslrun.Vec([f32(1 / nx), f32(1 / ny)]),
                       gridSize=glslrun.Vec([nx, ny], "i"), tau=f32(tau), U0=f32(u0))
            outs = glslrun.run_pass(step, nx, ny, uni)
            tex = [outs["outA"], outs["outB"], outs["outC"]]
            nsteps += 1
            print(f"  step {nsteps} ({time.time() - t0:.1f}s)", flush=True)
        I.globals.get("macro").a[:] = tex[2].reshape(-1)          # readMacro()
        I.call("updateFieldsFromMacro")
        row = dict(frame=frame, maxS=I.globals.get("maxS"), cpMin=I.globals.get("cpMin"),
                   cpMax=I.globals.get("cpMax"))
        if frame % forces_every == 0:
            I.call("computeForces")
            row.update(CLsmooth=I.globals.get("CLsmooth"), CDsmooth=I.globals.get("CDsmooth"),
                       sepFrac=I.globals.get("sepFrac"))
        series.append(row)
    F = np.concatenate([np.moveaxis(tex[0], 2, 0), np.moveaxis(tex[1], 2, 0), tex[2][None, :, :, 0]])
    fields = dict(F=F, rho=tex[2][:, :, 1], ux=tex[2][:, :, 2], uy=tex[2][:, :, 3],
                  U=I.globals.get("Ufield").a.reshape(ny, nx).copy(),
                  V=I.globals.get("Vfield").a.reshape(ny, nx).copy(),
                  Cp=I.globals.get("CpField").a.reshape(ny, nx).copy())
    stats = (I.globals.get("maxS"), I.globals.get("cpMin"), I.globals.get("cpMax"))
    return init_vals, series, fields, tex, stats


def to_unorm8(col):
    """RGBA8 default framebuffer (OpenGL ES 3.0, 2.1.6.1): clamp to [0, 1], scale by 255, round to nearest."""
    return np.floor(np.clip(col, 0, 1).astype(np.float32) * np.float32(255) + np.float32(0.5)).astype(np.uint8)


def render(page, nx, ny, u0, mask, texC, stats):
    sh = glslrun.Shader(extract.shader(page, "RENDER_FS_SRC"))
    mask_tex = glslrun.Sampler((mask.astype(np.float32) / np.float32(255.0)).reshape(ny, nx, 1))
    out = {}
    for mode in (0, 1, 2):
        uni = dict(texC=glslrun.Sampler(texC), texMask=mask_tex, texel=glslrun.Vec([f32(1 / nx), f32(1 / ny)]),
                   fieldMode=mode, U0=f32(u0), maxS=f32(stats[0]), cpMin=f32(stats[1]), cpMax=f32(stats[2]),
                   vortScale=f32(0.06))
        col = glslrun.run_pass(sh, nx, ny, uni)["fragColor"]
        out[mode] = to_unorm8(col)
        print("render mode", mode, flush=True)
    return out


def advect_pins(page, arrays):
    """sampleScalar/sampleUV/advect (HTML:616-639, 754-767) on tunnel A's final Ufield/Vfield."""
    from refexec.jsrun import UNDEF, TypedArray
    nx, ny = 48, 24
    I = js_tunnel(page, nx, ny)
    for fn in ("sampleScalar", "sampleUV", "advect"):
        I.run(extract.js_function(page, fn))
    IN = TypedArray("Uint8Array", nx * ny)
    IN.a[:] = arrays["tunA_mask"].reshape(-1)
    I.globals.set("sol", {"IN": IN})
    I.globals.get("Ufield").a[:] = arrays["tunA_U"].reshape(-1)
    I.globals.get("Vfield").a[:] = arrays["tunA_V"].reshape(-1)
    rng = np.random.default_rng(3)
    pts = np.column_stack([rng.uniform(-0.45, 1.45, 400), rng.uniform(-0.5, 0.5, 400)])
    pts[:40, 0] = rng.uniform(0.0, 1.0, 40)
    pts[:40, 1] = rng.uniform(-0.08, 0.08, 40)
    arrays["advect_pts"] = pts
    for key, dt in (("advect_dt16", 16.0), ("advect_dt2000", 2000.0)):     # the second hits the maxDisp clamp
        out = np.full((400, 4), np.nan)
        for k, (x, y) in enumerate(pts):
            r = I.call("advect", {"x": float(x), "y": float(y)}, dt)
            if r is not None and r is not UNDEF:
                out[k] = [r["nx"], r["ny"], r["speed"], 1.0]
            else:
                out[k, 3] = 0.0
        arrays[key] = out
    return dict(tunnel="A", dts=[16.0, 2000.0])


def main():
    page = extract.read_page()
    arrays = {}
    geom_cases = [("naca0012", 5.0, 320, 160), ("naca2412", 6.0, 320, 160), ("naca4412", 10.0, 320, 160),
                  ("naca6409", -7.5, 320, 160), ("clark_y", 6.0, 320, 160), ("naca2412", 25.0, 100, 37),
                  ("naca2412", 5.0, 333, 171), ("clark_y", 6.0, 2048, 1024), ("naca0012", 0.0, 2048, 1024),
                  ("naca4412", 10.0, 4096, 2048)]
    geom, garr = geometry(page, geom_cases)
    arrays.update(garr)
    gold = json.load(open(os.path.join(HERE, "golden.json")))
    user = gold["parser"]["naca0012_selig_test_main"]["coords"]
    user = [[round(float(x), 6), round(float(y), 6)] for x, y in user]
    user_geom = [user_coords_geometry(page, user, a, nx, ny) for a, nx, ny in ((0.0, 2048, 1024), (6.0, 320, 160))]

    tunnels = []
    # case A: the page's own pipeline on a small lattice -- NACA 2412 at alpha = 8, page defaults
    I = js_tunnel(page, 48, 24)
    pan = I.call("panelise", I.call("rotate", I.globals.get("SHAPES")["naca2412"](), 8.0))
    maskA = I.call("rasterMask", pan["xp"], pan["yp"]).a.reshape(24, 48).copy()
    # case B: faster inlet, lower tau, solids touching every border and the outlet column
    rng = np.random.default_rng(17)
    maskB = (rng.random((20, 40)) < 0.07).astype(np.uint8) * 255
    maskB[0, 4:9] = 255; maskB[-1, 20:26] = 255; maskB[6:9, 0] = 255; maskB[11:15, -1] = 255; maskB[3:6, -2] = 255
    # case C: broadside plate, fast inlet, tau close to 1/2 -- the rho / |u| clamps fire (HTML:340-350)
    maskC = np.zeros((18, 36), np.uint8)
    maskC[3:15, 12:14] = 255
    for name, nx, ny, u0, tau, mask, frames in (("A", 48, 24, 0.06, 0.58, maskA, 6), ("B", 40, 20, 0.1, 0.52, maskB, 3),
                                                ("C", 36, 18, 0.25, 0.505, maskC, 4)):
        print("tunnel", name, flush=True)
        init_vals, series, fields, tex, stats = run_tunnel(page, nx, ny, u0, tau, mask, frames, 4, 3)
        arrays[f"tun{name}_mask"] = mask
        arrays[f"tun{name}_init"] = init_vals
        for k, v in fields.items():
            arrays[f"tun{name}_{k}"] = v
        imgs = render(page, nx, ny, u0, mask, tex[2], stats)
        for mode, img in imgs.items():
            arrays[f"tun{name}_rgba{mode}"] = img
        tunnels.append(dict(name=name, nx=nx, ny=ny, u0=u0, tau=tau, frames=frames, steps_per_frame=4,
                            forces_every=3, series=series, final_stats=list(stats)))

    advect_info = advect_pins(page, arrays)

    with open(os.path.join(HERE, "ref_pins.json"), "w") as fh:
        json.dump(dict(source="pages/airfoil_flow_lbm_aerolab.html executed by tests/refexec (minimal JS/GLSL interpreters)",
                       geometry=geom, user_coords_geometry=user_geom, tunnels=tunnels, advect=advect_info), fh, indent=1)
    np.savez_compressed(os.path.join(HERE, "ref_pins.npz"), **arrays)
    print("wrote ref_pins.json / ref_pins.npz")


if __name__ == "__main__":
    main()
