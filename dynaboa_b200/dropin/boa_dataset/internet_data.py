from dynaboa_b200.datasets import Internet_dataset  # noqa: F401
