class LightningLoggerBase:
    pass
