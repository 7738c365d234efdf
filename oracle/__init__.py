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

PARITY UNPINNED: the reference ships no test, golden vector or known-answer
fixture for the LBM path and its implementation (GLSL ES 3.00 in a browser)
cannot be executed in the build container (no JS engine, no browser).  The
oracle is therefore a careful restatement, cross-checked by a second,
independently written NumPy restatement (``oracle/lbm_numpy.py``) that must
agree bitwise with the C one (``oracle/lbm_ref.c``), and by the surveyor's
probe pins recorded in SURVEY.md section 8(c).
"""
