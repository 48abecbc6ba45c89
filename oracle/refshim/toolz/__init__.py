"""Stand-in for toolz (absent from this image): summary_utils.py:4 imports sliding_window."""


def sliding_window(n, seq):
    seq = list(seq)
    return [tuple(seq[i:i + n]) for i in range(len(seq) - n + 1)]
