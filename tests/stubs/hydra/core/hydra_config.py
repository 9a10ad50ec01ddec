class HydraConfig:
    @staticmethod
    def get():
        raise RuntimeError("no Hydra runtime in the test stand-in")
