"""CPU oracle for the box-attention hot path.  TEST INFRASTRUCTURE ONLY.

Nothing under ``oracle/`` is part of the product.  Only ``tests/``,
``__graft_entry__.smoke()`` and the ``cpu_baseline`` / ``--impl reference``
legs of ``bench.py`` may import it, and only as the checker (or as the timed
CPU baseline) -- never as the thing shipped.  ``boxer_b200`` must not import
this package; ``tests/test_cpu_host.py::test_product_never_imports_the_oracle`` enforces that.

Two independent restatements live here:

* ``plain.py``   -- the reference's own oracle formulation (per-level
  ``F.grid_sample`` + weighted sum), following
  ``/root/reference/tests/box_attn_test.py:9-42``,
  ``/root/reference/tests/instance_attn_test.py:11-63`` and
  ``/root/reference/e2edet/utils/general.py:289-324``.
* ``kernel_ref.c`` (+ ``kernel_ref.py`` ctypes binding) -- a plain-C
  restatement of the reference CUDA kernels' arithmetic (pixel = loc*size-0.5,
  window test, per-corner zero padding, gradient formulas), following
  ``/root/reference/e2edet/module/ops/src/box_attn/box_attn_kernel.cuh:34-184,274-349``
  and ``.../instance_attn/instance_attn_kernel.cuh:98-187,282-364``.

Parity pinning: the reference ships no golden vectors.  ``tests/golden/*.npz``
were produced by executing the reference's *own* Python oracle functions (text
taken from ``/root/reference`` at generation time by
``tests/golden/make_golden.py``) on the reference tests' seeded inputs; both
restatements are checked against those fixtures in ``tests/test_oracle.py``.
"""
