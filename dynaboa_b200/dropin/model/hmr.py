from dynaboa_b200.hmr import *  # noqa: F401,F403
from dynaboa_b200.hmr import hmr, HMR  # noqa: F401
