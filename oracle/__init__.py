"""oracle/ -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

CPU restatement (numpy / scipy.sparse / torch-CPU fp32) of the reference's
algorithm for the one hot path this repo accelerates: the Hyper-SAGNN
hyperedge scorer over (cell, bin_i, bin_j) triplets and the imputation driver
around it (reference: higashi/Higashi_backend/Modules.py, higashi/Impute.py,
higashi/Higashi_wrapper.py; every function cites the file:line it follows).

Who may import this package: `tests/`, `__graft_entry__.smoke()` and the
`cpu_baseline` / `--impl reference` legs of `bench.py` -- as the checker or the
reported CPU baseline only.  The product (`higashi_b200/`) never imports it and
has no CPU fallback.

Parity status: PINNED.  The reference has no tests or golden vectors of its own
(SURVEY.md section 4), so the oracle is pinned against outputs of the
unmodified reference run in the build container: `tests/golden/make_golden.py`
generated `tests/golden/*.npz`, and `tests/test_oracle_golden.py` checks every
oracle function against them (`-m "not gpu"`).

Third-party arithmetic the reference delegates to (not vendored under
/root/reference; versions are those of this image because the reference's own
pins -- scipy==1.7.3, pandas==1.3.4 in setup.py:13-14 -- are not installable
here): torch 2.11 CPU kernels, scipy.sparse 1.18 CSR add/scale,
scipy.stats.norm.pdf, sklearn 1.9 `normalize` / `pairwise_distances`, numpy 2.3
sort/argsort/quantile.  Their published definitions are restated where used.
"""
