"""CPU oracle for the SGFE solve hot path of ExtendableASGFEM.jl (v1.0.1).

TEST INFRASTRUCTURE ONLY.  Nothing in the product path (the package
`extendableasgfem.jl_b200`, `libasgfem_cuda.so`) may import, call or link
anything in this directory; only `tests/`, `__graft_entry__.smoke()` and the
`cpu_baseline` / `--impl reference` legs of `bench.py` do, and only as the checker.

The reference is pure Julia and neither `julia` nor its package depot exists in
this image, so this is a restatement in numpy/scipy (+ one C file
for the timed CPU baseline) that follows the cited reference lines one by one.

Parity pins
-----------
* polynomials / ONBasis / TensorizedBasis triple products: PINNED against the
  known answers of the reference's own test/runtests.jl:33-127 (tests/test_oracle_*).
* everything else (G, neighbour tables, multi-index management, coefficient,
  assembly, operator, preconditioner, Krylov, estimator): PARITY UNPINNED by the
  reference's tests (it has none for those); pinned only by in-repo identities
  (G == quadrature triple_product_y, mul! == assembled block matrix of
  solve_full_primal!, PCG == GMRES == direct solve).  Third-party semantics that
  are assumed are listed in SURVEY.md Appendix B and isolated in `mesh.py`/`fem.py`.
"""
