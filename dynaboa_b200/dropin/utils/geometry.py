from dynaboa_b200.geometry import (batch_rodrigues, rot6d_to_rotmat, perspective_projection,  # noqa: F401
                                   rotation_matrix_to_angle_axis, project_normalized)
