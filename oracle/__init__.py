"""CPU oracle for the dqgp hot path — TEST INFRASTRUCTURE, NOT PRODUCT.

Only ``tests/``, ``__graft_entry__.smoke()`` and the ``cpu_baseline`` / ``--impl reference`` legs of
``bench.py`` may import this package, and only as the checker or the timed CPU baseline.  The
product path (``dqgp_b200``) never imports it and fails loudly when the CUDA library is missing.

What it restates (NumPy/SciPy, fp64 / complex128), with the reference file:line each part follows:

* ``circuits``     the four encoding circuits the reference constructs from ``squlearn==0.9.1``
                   (ctor sites ``main.py:68-83``, ``agent_riemannian.py:51-66``) — the third-party
                   source is NOT in this container, so these gate lists are restated from the
                   published squlearn 0.9.1 circuit library and marked [UPSTREAM-RECALLED].
* ``statevector``  exact statevector simulation (two independent implementations) + Pauli features.
* ``qkernels``     FidelityKernel / ProjectedQuantumKernel duck-types (``main.py:118-137``) and the
                   sklearn outer kernels (verified against the installed scikit-learn in tests).
* ``torus``        ``riemannian_optimizer.py:26-51,73-129,302-368``.
* ``agent_step``   ``agent_riemannian.py:209-277`` (job list, central difference) and ``:410-486``.
* ``driver``       ``main.py:2403-2555`` (ADMM loop), ``:1364-1488`` (prediction), ``:1546-1552`` (NLPD).

PARITY STATUS
-------------
* Pinned against the *real* reference code run in this container: everything in
  ``riemannian_optimizer.py`` (imported unmodified) and the whole of ``RiemannianAgent.train_and_update``
  / ``main.predict_quantum_gp`` / ``main.main`` executed unmodified with the squlearn import satisfied by
  ``oracle.fake_squlearn`` (our oracle kernels plugged in exactly at the squlearn boundary).  The
  golden vectors live in ``tests/golden/*.npz|json`` and are produced by ``tests/golden/make_golden.py``.
* Pinned against installed scikit-learn 1.9.0: the Gaussian / Matern / ExpSineSquared outer kernels.
* **PARITY UNPINNED** at the squlearn boundary itself: circuit gate lists, qubit ordering and the
  statevector conventions are restated from memory of squlearn 0.9.1 / Qiskit and could not be checked
  against the real package (no wheel, no network).  What pins them here: two independent simulators,
  analytic known-answer tests and invariants (``tests/test_oracle_circuits.py``).
"""
