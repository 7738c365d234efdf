#!/usr/bin/env python
"""More golden vectors from the reference's own source (second batch, see make_ref_pins.py).

    python tests/golden/make_ref_pins2.py            # needs /root/reference; writes ref_pins2.{json,npz}

Four small tunnels that the first batch does not cover (round-1 verdict, item 8): relaxation times
other than the page's 0.58, slider changes in the middle of a run (U0 every frame, tau between
frames -- the shader takes both as uniforms, HTML:228-231, 520-521), open-trailing-edge user
coordinates and the Clark-Y table rasterised by the page's own pipeline on odd lattice sizes.
The step shader is executed per fragment by tests/refexec/glslrun.py, the statistics and force code
by tests/refexec/jsrun.py; only numeric outputs are stored.
"""
import json
import os
import sys
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import make_ref_pins as M  # noqa: E402
from refexec import extract, glslrun  # noqa: E402
from refexec.glslrun import f32  # noqa: E402
from refexec.jsrun import TypedArray  # noqa: E402


def run_tunnel_controls(page, nx, ny, mask, controls, steps_per_frame, forces_every):
    """The page's frame loop with per-frame slider values: controls = [(u0, tau), ...]."""
    I = M.js_tunnel(page, nx, ny)
    u0_0 = controls[0][0]
    I.globals.set("U0", float(u0_0))
    init = I.call("equilibriumInitData", float(u0_0))
    tex = [init[k].a.reshape(ny, nx, 4).copy() for k in ("dA", "dB", "dC")]
    step = glslrun.Shader(extract.shader(page, "STEP_FS_SRC"))
    IN = TypedArray("Uint8Array", nx * ny)
    IN.a[:] = mask.reshape(-1)
    I.globals.set("sol", {"IN": IN})
    mask_tex = glslrun.Sampler((mask.astype(np.float32) / np.float32(255.0)).reshape(ny, nx, 1))
    series = []
    for frame, (u0, tau) in enumerate(controls, start=1):
        I.globals.set("U0", float(u0))                      # the slider handler, HTML:956-959
        for _ in range(steps_per_frame):
            uni = dict(texA=glslrun.Sampler(tex[0]), texB=glslrun.Sampler(tex[1]), texC=glslrun.Sampler(tex[2]),
                       texMask=mask_tex, texel=glslrun.Vec([f32(1 / nx), f32(1 / ny)]),
                       gridSize=glslrun.Vec([nx, ny], "i"), tau=f32(tau), U0=f32(u0))
            outs = glslrun.run_pass(step, nx, ny, uni)
            tex = [outs["outA"], outs["outB"], outs["outC"]]
        I.globals.get("macro").a[:] = tex[2].reshape(-1)          # readMacro()
        I.call("updateFieldsFromMacro")
        row = dict(frame=frame, u0=u0, tau=tau, maxS=I.globals.get("maxS"), cpMin=I.globals.get("cpMin"),
                   cpMax=I.globals.get("cpMax"))
        if frame % forces_every == 0:
            I.call("computeForces")
            row.update(CLsmooth=I.globals.get("CLsmooth"), CDsmooth=I.globals.get("CDsmooth"),
                       sepFrac=I.globals.get("sepFrac"))
        series.append(row)
        print("  frame", frame, flush=True)
    F = np.concatenate([np.moveaxis(tex[0], 2, 0), np.moveaxis(tex[1], 2, 0), tex[2][None, :, :, 0]])
    return series, dict(F=F, rho=tex[2][:, :, 1], ux=tex[2][:, :, 2], uy=tex[2][:, :, 3])


def page_mask(page, nx, ny, coords, alpha):
    I = M.js_tunnel(page, nx, ny)
    pan = I.call("panelise", I.call("rotate", coords, float(alpha)))
    return I.call("rasterMask", pan["xp"], pan["yp"]).a.reshape(ny, nx).copy()


def main():
    page = extract.read_page()
    gold = json.load(open(os.path.join(HERE, "golden.json")))
    user = [[round(float(x), 6), round(float(y), 6)] for x, y in gold["parser"]["naca0012_selig_test_main"]["coords"]]
    I0 = M.js_tunnel(page, 8, 8)
    clark = I0.globals.get("SHAPES")["clark_y"]()
    n6409 = I0.globals.get("SHAPES")["naca6409"]()
    rng = np.random.default_rng(29)
    maskF = (rng.random((17, 33)) < 0.06).astype(np.uint8) * 255
    maskF[0, 3:6] = 255; maskF[-1, 9:12] = 255; maskF[5:8, 0] = 255; maskF[9:12, -1] = 255
    cases = [
        # open-TE user coordinates (13 points) at alpha = 7 on an odd lattice; U0 and tau change mid-run
        ("D", 37, 23, page_mask(page, 37, 23, user, 7.0),
         [(0.05, 0.7), (0.05, 0.7), (0.08, 0.7), (0.08, 0.62), (0.08, 0.62), (0.06, 0.62)]),
        # the Clark-Y table (open TE), strongly over-relaxed towards tau = 0.9
        ("E", 45, 19, page_mask(page, 45, 19, clark, 12.0), [(0.09, 0.9)] * 3),
        # tau > 1 (under-relaxation), slow inlet, solids on every border
        ("F", 33, 17, maskF, [(0.03, 1.3)] * 3),
        # the U0 slider moved on every frame across its whole range (HTML:41)
        ("G", 64, 32, page_mask(page, 64, 32, n6409, -7.5), [(0.03, 0.58), (0.05, 0.58), (0.07, 0.58), (0.1, 0.58), (0.044, 0.58), (0.06, 0.58)]),
    ]
    arrays, tunnels = {}, []
    for name, nx, ny, mask, controls in cases:
        t0 = time.time()
        print("tunnel", name, nx, ny, flush=True)
        series, fields = run_tunnel_controls(page, nx, ny, mask, controls, 4, 3)
        arrays[f"tun{name}_mask"] = mask
        for k, v in fields.items():
            arrays[f"tun{name}_{k}"] = v
        tunnels.append(dict(name=name, nx=nx, ny=ny, controls=controls, steps_per_frame=4, forces_every=3, series=series))
        print(f"  {time.time() - t0:.0f} s", flush=True)
    with open(os.path.join(HERE, "ref_pins2.json"), "w") as fh:
        json.dump(dict(source="pages/airfoil_flow_lbm_aerolab.html executed by tests/refexec (second batch)", tunnels=tunnels), fh, indent=1)
    np.savez_compressed(os.path.join(HERE, "ref_pins2.npz"), **arrays)
    print("wrote ref_pins2.json / ref_pins2.npz")


if __name__ == "__main__":
    main()
