"""pioneer.envs -> pioneer_b200.envs"""
