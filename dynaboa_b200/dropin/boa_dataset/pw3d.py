from dynaboa_b200.datasets import PW3D  # noqa: F401
