"""CPU oracle for the ppx learner hot path  --  TEST INFRASTRUCTURE, NOT PRODUCT.

This package restates, in plain numpy / torch-CPU, the algorithms of the reference
(BoogaQ/PPO-exploration: buffer.py, algorithms.py, models.py, util.py, evolution_strategies.py,
sil_module.py) that the CUDA library in ``ppo-exploration_b200/csrc`` replaces.  Every function
cites the reference file:line it follows.

Rules (enforced by tests/test_cabi_and_host.py::test_product_never_touches_the_oracle_or_cpu_fallbacks):
  * only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
    ``--impl reference`` legs may import this package;
  * the product package never imports it and has no CPU fallback.

Parity pinning: the reference ships NO tests, golden vectors or known-answer tests (SURVEY.md §4,
§8c).  The oracle is therefore pinned against outputs of the *unmodified reference itself*, run in
the build container through ``tests/golden/ref_shim.py`` and committed as ``tests/golden/*.npz``
by ``tests/golden/make_golden.py``.  ``tests/test_oracle_golden.py`` checks every oracle function
against those fixtures (bit-exact for integer work and for the sequential float recurrences).
Third-party arithmetic the reference leans on (numpy MT19937 ``permutation``/``randn``, torch CPU
autograd / Adam / clip_grad_norm_, sklearn NearestNeighbors) is unpinned upstream; parity is defined
against the versions in this image (numpy 2.3.5, torch 2.11.0, scikit-learn 1.9.0).
"""
