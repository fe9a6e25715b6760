class LBFWrapper:
    def __init__(self, *a, **k):
        raise RuntimeError("lbforaging is not installed; only env_type=pz simple_spread_v3 is available in the stub")
