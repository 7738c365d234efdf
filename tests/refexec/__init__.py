"""Minimal JavaScript / GLSL interpreters that execute the reference's own source text.

TEST INFRASTRUCTURE ONLY.  Used by tests/golden/make_ref_pins.py (in the build container, where
/root/reference exists) to produce golden vectors FROM THE REFERENCE ITSELF, against which the
CPU oracle is pinned (tests/test_reference_pins.py).  Nothing here is imported by the product.
"""
