"""Stand-in for omegaconf (criterions/label_smoothed_cross_entropy.py:11 uses only II)."""
def II(key):  # interpolation marker; the harness passes sentence_avg explicitly
    return False
