def __getattr__(name):
    raise RuntimeError("matplotlib stand-in: plotting is not available")
