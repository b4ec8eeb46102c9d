DEFAULT_CONFIG = {}


class PPO:  # placeholder, replaced below in this round
    pass
