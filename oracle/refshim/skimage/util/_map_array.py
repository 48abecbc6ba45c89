"""skimage.util.map_array as the reference uses it (csr.py:133-144): integer lookup of every element of `input_arr` in
`input_vals`, returning the matching `output_vals`, 0 where absent.  scikit-image does this with a Cython hash map;
here a binary search in the (already ascending, csr.py:131) keys.  The time spent in here is reported next to the
CPU baseline (bench.py `cpu_baseline.map_array_share`)."""
import time

import numpy as np

SECONDS = [0.0]   # accumulated wall time inside map_array (bench.py reads and resets it)


def map_array(input_arr, input_vals, output_vals, out=None):
    t0 = time.perf_counter()
    input_arr = np.asarray(input_arr)
    input_vals = np.asarray(input_vals)
    output_vals = np.asarray(output_vals)
    if input_vals.size > 1 and not np.all(input_vals[1:] >= input_vals[:-1]):
        order = np.argsort(input_vals, kind='stable')
        input_vals, output_vals = input_vals[order], output_vals[order]
    flat = input_arr.ravel()
    res = np.zeros(flat.size, dtype=output_vals.dtype) if out is None else out.reshape(-1)
    if input_vals.size:
        pos = np.searchsorted(input_vals, flat)
        np.minimum(pos, input_vals.size - 1, out=pos)
        hit = input_vals[pos] == flat
        res[...] = np.where(hit, output_vals[pos], 0)
    SECONDS[0] += time.perf_counter() - t0
    return res.reshape(input_arr.shape)


class ArrayMap:
    def __init__(self, in_values, out_values):
        self.in_values = np.asarray(in_values)
        self.out_values = np.asarray(out_values)

    def __len__(self):
        return len(self.in_values)

    def __getitem__(self, index):
        scalar = np.isscalar(index)
        out = map_array(np.atleast_1d(np.asarray(index)), self.in_values, self.out_values)
        return out[0] if scalar else out
