from dynaboa_b200.maml import MAML  # noqa: F401
