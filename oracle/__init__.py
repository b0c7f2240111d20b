"""CPU oracle for the splatting hot path — TEST INFRASTRUCTURE ONLY.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs may import this package.  The product
(``gaussiansplattingmlx_b200`` → ``libgsb.so``) never imports, links, calls or falls back to it.

Two implementations sit behind one numpy-level interface (``oracle.port`` / ``oracle.ref``):

* ``oracle/gsb_oracle.c`` — a plain-C restatement ("port") of the reference's kernels, each
  function citing the reference file:line it follows.  Always available (gcc).
* ``oracle/_ref/libgsref.so`` — the reference's OWN shipped kernel source
  (``GaussianSplattingMlx/Slang/*_mlx.json``) compiled for the CPU by ``oracle/build_ref.py``
  where it lies under ``/root/reference``.  Used to pin the port (tests/test_oracle_pinning.py)
  and, when present, as the CPU baseline of kind "reference".

Parity status: PINNED — the port is checked against the compiled reference kernels on seeded
scenes (bit-exact for the forward projection / binning / raster, <=1e-5 relative for the
backward passes) and against the reference's own known-answer unit tests (SH polynomial,
build_rotation, build_scaling_rotation).  Adam is the exception: it lives in the un-vendored
mlx-swift 0.30.6 dependency and no reference test pins its arithmetic ("parity unpinned" for
that one function; restated from the published update rule).
"""
