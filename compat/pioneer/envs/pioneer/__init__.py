"""pioneer.envs.pioneer (reference: pioneer/envs/pioneer/__init__.py:1) -> pioneer_b200.envs.pioneer"""
from pioneer_b200.envs.pioneer import PioneerKinematicConfig, PioneerKinematicEnv  # noqa: F401
