class GlobalHydra:
    _inst = None

    @classmethod
    def instance(cls):
        if cls._inst is None:
            cls._inst = cls()
        return cls._inst

    def clear(self):
        import hydra
        hydra._state["config_dir"] = None

    def is_initialized(self):
        import hydra
        return hydra._state["config_dir"] is not None
