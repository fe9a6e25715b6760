"""Duck-type every env adapter follows (reference ``cleanmarl/env/common_interface.py:5-23``)."""


class CommonInterface:
    _METHODS = ("step", "reset", "get_avail_actions", "get_action_size", "get_state",
                "get_state_size", "get_obs_size", "close", "sample")

    def __getattr__(self, name):
        if name in CommonInterface._METHODS:
            raise NotImplementedError(name)
        raise AttributeError(name)
