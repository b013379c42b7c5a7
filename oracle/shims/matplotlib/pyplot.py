def __getattr__(name):  # any plt.xxx(...) is a no-op
    def _noop(*a, **k):
        return None
    return _noop
