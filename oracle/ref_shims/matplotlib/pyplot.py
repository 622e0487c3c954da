def __getattr__(name):
    raise RuntimeError("matplotlib shim: plotting is not available (%s)" % name)
