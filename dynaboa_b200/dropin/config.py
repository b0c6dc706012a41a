from dynaboa_b200.config import *  # noqa: F401,F403
