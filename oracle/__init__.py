"""CPU oracle (TEST INFRASTRUCTURE ONLY).

float64 restatement of the reference's matrix-free Laplace hot path
(/root/reference/src/{ggn,stochtrace,sample,lla,matfree_monkeypatch,utils}.py and the
third-party matfree / jax.scipy.sparse.linalg.cg algorithms those files call).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import this package.  The product (laplace-inducing-points_b200/) never does.
"""
