"""Checks of our own in place of compute-sanitizer (closed on the GPU pool): guard gaps behind every allocation of the
native runners, arenas pre-filled with a byte pattern.  An out-of-bounds write damages a gap; a result that changes with the
pattern was computed from uninitialised memory; a result that changes between identical runs points at a race."""
import numpy as np
import torch

from skan_b200 import _native, _pipeline


class guarded:
    def __init__(self, fill):
        self.fill = fill

    def __enter__(self):
        self.lib = _native.library()
        self.lib.call('skb_set_option', b'arena_guard', 1)
        _pipeline.DEBUG_FILL = self.fill
        _pipeline._arena_cache.clear()
        return self

    def __exit__(self, *exc):
        self.lib.call('skb_set_option', b'arena_guard', 0)
        _pipeline.DEBUG_FILL = None
        _pipeline.LAST_ARENAS = None
        _pipeline._arena_cache.clear()
        return False

    def check_gaps(self):
        """every guard gap of the last native call still holds the fill byte; returns the number of gaps checked"""
        total = 0
        for which, arena in enumerate(_pipeline.LAST_ARENAS):
            offs = np.zeros(512, dtype=np.int64)
            n = self.lib._dll.skb_debug_guard_offsets(which, offs.ctypes.data, 512)
            assert 0 < n <= 512
            for o in offs[:n]:
                if o < 0:       # space that an earlier, rewound allocation of the same call used: not checkable
                    continue
                gap = arena[int(o): int(o) + 256]
                assert gap.numel() == 256 and bool((gap == self.fill).all()), \
                    f'guard gap at offset {int(o)} of arena {which} was written to'
            total += int(n)
        return total
