from ...config import PioneerKinematicConfig
from ...batched_env import BatchedPioneerEnv
from .pioneer_knm_env import PioneerKinematicEnv
