"""CPU oracle for the sdf-nmpc per-iteration NMPC hot path.

TEST INFRASTRUCTURE ONLY.  This package is a float64 numpy restatement of the
reference algorithm (ntnu-arl/sdf-nmpc, acados 0.3.1 + L4CasADi 1.4.1
semantics).  It is the *checker* for the CUDA path: only ``tests/``,
``__graft_entry__.smoke()`` and the CPU-baseline / ``--impl reference`` legs of
``bench.py`` may import it.  Nothing under ``sdf_nmpc_b200/`` imports it and the
product path has no CPU fallback.

Parity status (see DESIGN.md "Oracle"):

* ``oracle.sdf``     - PINNED against the reference's own torch modules
  (``sdf_nmpc/network/neural_df.py``, ``sdf_nmpc/utils/embeddings.py``) through
  the committed fixtures ``tests/golden/sdf_*.npz`` (generator:
  ``tests/golden/make_golden.py``, run where ``/root/reference`` exists).
* ``oracle.encoder`` - PINNED the same way against ``sdf_nmpc/network/vae.py``
  + ``sdf_nmpc/utils/preprocessing.py`` (``tests/golden/encoder_*.npz``).
* ``oracle.model_att`` - values PINNED against the reference's numpy branches
  of ``sdf_nmpc/utils/math.py`` composed as ``model/quad_rollpitchyawrate.py``
  does (``tests/golden/model_att.npz``); Jacobians pinned by finite differences
  and torch.float64 autograd of the literal expressions.
* ``oracle.qp`` / ``oracle.rti`` - **parity unpinned**: acados 0.3.1, HPIPM and
  CasADi are third-party dependencies whose source is not in the reference
  tree and which cannot be installed here; the reference ships no golden
  vectors.  The QP semantics follow SURVEY.md section 8a-5 (from memory of the
  public acados source) and are cross-checked only against scipy on small
  problems and against KKT residuals.
"""
