"""CPU oracle for the particle-SCP solve path.  TEST INFRASTRUCTURE ONLY.

This package restates, in NumPy/SciPy, the algorithm of the reference
(StanfordASL/pmpc) for the one hot path this repository accelerates:

    pmpc.solve -> static_backend.aff_solve -> c_lqp_solve -> PMPC.jl lqp_solve
               -> lqp_repr_Pq / lqp_repr_Ab / lqp_repr_Gla -> OSQP

Only ``tests/``, ``__graft_entry__.smoke()`` and the ``cpu_baseline`` /
``--impl reference`` legs of ``bench.py`` may import it; the product package
``pmpc_b200`` never does (it fails loudly when the CUDA library is missing).

Parity status: **parity unpinned** for the convex sub-problem itself — the
reference holds no golden vectors for it and neither Julia nor OSQP exist in the
build image.  What *is* pinned:

* the Dubins linearisation (``oracle.dubins``) against the reference's own
  ``tests/dubins_car.py`` run here (fixtures in ``tests/golden/``);
* the SCP outer loop (``oracle.scp``) against the reference's real
  ``pmpc.scp_mpc.scp_solve`` executed here with ``aff_solve`` routed to
  ``oracle.lqp`` (fixtures in ``tests/golden/``);
* the QP (``oracle.lqp``) through a solver-independent KKT certificate of the
  strictly convex problem built exactly as ``PMPC.jl/src/lqp_utils.jl`` does.
"""
