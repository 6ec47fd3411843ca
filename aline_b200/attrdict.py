"""``AttrDict``: the reference uses the ``attrdictionary`` package (tasks/base_task.py:6, model/base.py).
Use it when installed so that objects interoperate; otherwise a minimal stand-in."""
try:  # pragma: no cover - depends on the environment
    from attrdictionary import AttrDict  # type: ignore
except Exception:  # noqa: BLE001
    class AttrDict(dict):
        """dict with attribute access (the subset of attrdictionary.AttrDict the ALINE code relies on)."""

        def __getattr__(self, k):
            try:
                return self[k]
            except KeyError:
                raise AttributeError(k) from None

        def __setattr__(self, k, v):
            self[k] = v

        def __delattr__(self, k):
            try:
                del self[k]
            except KeyError:
                raise AttributeError(k) from None
