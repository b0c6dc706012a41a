from dynaboa_b200.hmr import hmr, HMR  # noqa: F401
from dynaboa_b200.smpl import SMPL  # noqa: F401
