from dynaboa_b200.dataprocess import crop, get_transform, j2d_processing, process_sample, transform  # noqa: F401
