"""Minimal stand-in for `scooby` (version reporter; not on the hot path)."""


class Report:
    def __init__(self, *args, **kwargs):
        pass
