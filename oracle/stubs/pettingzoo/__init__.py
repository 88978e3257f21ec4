"""Import stub (test infrastructure): SCS_Game only inherits from AECEnv."""


class AECEnv:
    pass
