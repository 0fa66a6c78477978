"""Minimal MLX-API shim over PyTorch-CPU -- TEST INFRASTRUCTURE ONLY (see oracle/README.md).

Lets the *unmodified* reference package (/root/reference/flux) be imported and executed in a
container that has no MLX wheel, so that golden vectors can be produced from the reference's own
code (oracle/gen_golden.py).  Only the MLX entry points the reference's Flux path touches exist.
Semantics follow MLX's public documentation; nothing here is shipped or measured.
"""
