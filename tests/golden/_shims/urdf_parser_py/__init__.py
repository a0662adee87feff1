"""Minimal stand-in for the third-party `urdf_parser_py` package (not installed in this
image, no network).  TEST INFRASTRUCTURE ONLY: it exists so that tests/golden/make_golden.py
can import the unmodified reference from /root/reference and record golden vectors.
It exposes just what the reference touches (models/utils.py:184-252, robots.py:6,24-53)."""
