"""Placeholder so `import uravu.relationship` does not fail (arrhenius.py is off the hot path)."""


class Relationship:  # pragma: no cover

    def __init__(self, *a, **k):
        raise NotImplementedError('uravu.relationship is not on the displacement-statistics path')
