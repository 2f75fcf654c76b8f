"""CPU oracle for the MSDeformAttn hot path.  TEST INFRASTRUCTURE ONLY.

Nothing under ``oracle/`` is part of the product.  Only ``tests/``,
``__graft_entry__.smoke()`` and the ``cpu_baseline`` / ``--impl reference`` legs of
``bench.py`` may import it, and there only as the checker or as the CPU arm being
timed next to the CUDA path.  The product package
(``category-agnostic-pose-estimation_b200``) never imports this package and fails
loudly when its CUDA library is missing.

Three independent restatements of the reference's
``ms_deform_attn_core_pytorch`` (``/root/reference/models/deformable_transformer.py:115-141``):

* :mod:`oracle.msda_numpy`   closed-form forward + backward in numpy (fp64 by default);
* :mod:`oracle.msda_torch`   the same per-level ``grid_sample`` formulation the reference
  calls, so it runs on the very ATen kernels (``grid_sampler_2d`` / ``_backward``)
  the reference's arithmetic lives in — this is what the CPU baseline times;
* ``oracle/msda_oracle.c``    plain-C (OpenMP) restatement for full-size checks,
  built by ``oracle/build_oracle.py`` into ``oracle/_build/``.

Parity pin: the reference's own tests hold no vectors for this path,
so all three are pinned against outputs of the reference itself, generated in the
build container by ``oracle/make_golden.py`` (imports ``/root/reference``) and
committed under ``tests/golden/``.
"""
