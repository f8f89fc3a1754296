"""CPU oracle for the particle-SCP solve path.  TEST INFRASTRUCTURE ONLY.

This package restates, in NumPy/SciPy, the algorithm of the reference
(StanfordASL/pmpc) for the one hot path this repository accelerates:

    pmpc.solve -> static_backend.aff_solve -> c_lqp_solve -> PMPC.jl lqp_solve
               -> lqp_repr_Pq / lqp_repr_Ab / lqp_repr_Gla -> OSQP

Only ``tests/``, ``__graft_entry__.smoke()`` and the ``cpu_baseline`` /
``--impl reference`` legs of ``bench.py`` may import it; the product package
``pmpc_b200`` never does (it fails loudly when the CUDA library is missing).

Parity status: **pinned on the reference's own printed output** — the reference's tests hold no golden vectors and
neither Julia nor OSQP / ECOS exist in the build image, but notebooks of the reference repository
keep the iteration tables its real solver stack (PMPC.jl behind ``pmpc.solve``) printed for fully specified,
deterministic problems; ``tests/test_reference_notebook_traces.py`` reproduces them row by row (``obj``, ``resid`` at the 4
printed digits): a hard-box single-particle run with ``oracle.lqp`` inside ``oracle.scp`` (and with the CUDA path), a
single-particle run with a slew penalty, one with zero control cost and a 20-particle consensus run through
``oracle.barrier`` (those runs smoothed
the box rows with the cone route's log barrier; same matrices, different treatment of the box rows).  The pin is as
sharp as the printed digits: 5e-4 relative, no full-precision vector of the reference's solver exists anywhere.
Beside it:

* the Dubins linearisation (``oracle.dubins``) against the reference's own
  ``tests/dubins_car.py`` run here (fixtures in ``tests/golden/``);
* the SCP outer loop (``oracle.scp``) against the reference's real
  ``pmpc.scp_mpc.scp_solve`` executed here with ``aff_solve`` routed to
  ``oracle.lqp`` (fixtures in ``tests/golden/``);
* the QP (``oracle.lqp``) through a solver-independent KKT certificate of the
  strictly convex problem built exactly as ``PMPC.jl/src/lqp_utils.jl`` does.
"""
