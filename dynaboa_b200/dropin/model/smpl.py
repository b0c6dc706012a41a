from dynaboa_b200.smpl import SMPL, SMPLOutput, get_smpl_faces, vertices2joints  # noqa: F401
