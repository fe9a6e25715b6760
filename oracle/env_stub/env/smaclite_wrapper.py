class SMACliteWrapper:
    def __init__(self, *a, **k):
        raise RuntimeError("smaclite is not installed; only env_type=pz simple_spread_v3 is available in the stub")
