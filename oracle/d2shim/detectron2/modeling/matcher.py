class Matcher(object):
    """constructed by ROIHeads.__init__ (center_heads.py:129-133) but only used in training."""

    def __init__(self, thresholds, labels, allow_low_quality_matches=False):
        self.thresholds, self.labels = thresholds, labels

    def __call__(self, *a, **k):
        raise RuntimeError("training-only symbol; not available in the oracle shim")
