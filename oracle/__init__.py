"""CPU oracle for the D2Q9 wind-tunnel hot path.  TEST INFRASTRUCTURE ONLY.

This package restates, on the CPU, the algorithm of the reference's
``pages/airfoil_flow_lbm_aerolab.html`` (geometry in float64, LBM step in
strict IEEE fp32, source operation order).  It exists so that the CUDA product
path in ``airfoil-cfd-tool_b200/`` can be checked against something.

Rules (see the task contract, section 3):

* only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s
  ``cpu_baseline`` / ``--impl reference`` legs may import or execute anything
  under ``oracle/``;
* the product package never imports it and has no CPU fallback.

PINNING STATUS.  The reference ships no test, golden vector or known-answer fixture for the LBM
path, and no JavaScript engine or browser exists in the build container, so the reference cannot be
run as a program.  Instead its OWN SOURCE TEXT (the JavaScript geometry / init / statistics / force /
particle-advection functions and the GLSL step and render shaders of
pages/airfoil_flow_lbm_aerolab.html) is executed by purpose-built minimal interpreters
(tests/refexec/, JavaScript with float64 semantics, GLSL with strict fp32 semantics) and the outputs
are committed as golden vectors (tests/golden/ref_pins.*, generator tests/golden/make_ref_pins.py).
The oracle reproduces every one of them exactly (tests/test_reference_pins.py): panel nodes and masks
of 12 geometry cases, the initial state, populations and rho/ux/uy after 24 and 12 steps on two small
lattices (bitwise), the autoscale statistics, the force EMAs and separation fraction, the
Ufield/Vfield/CpField arrays, the RGBA8 render output of all three field modes, and 800 particle
advections.  What this does NOT pin: the behaviour of a particular browser/GPU (GLSL ES `highp` does
not require correctly rounded division or forbid FMA contraction; V8's Math.* are fdlibm ports while
the interpreters and the oracle use glibc) -- the pinned semantics are "the reference's source,
evaluated with IEEE-754 operations in source order".  Additional cross-checks: a second,
independently written NumPy restatement (``oracle/lbm_numpy.py``) agrees bitwise with the C one
(``oracle/lbm_ref.c``), and the surveyor's probe values in SURVEY.md section 8(c) are reproduced.
"""
