from dynaboa_b200.prior import MaxMixturePrior  # noqa: F401
