from dynaboa_b200.pose_utils import compute_similarity_transform, compute_similarity_transform_batch  # noqa: F401
