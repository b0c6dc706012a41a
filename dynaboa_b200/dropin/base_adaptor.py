from dynaboa_b200.base_adaptor import BaseAdaptor  # noqa: F401
