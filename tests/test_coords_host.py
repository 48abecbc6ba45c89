"""skb_coords (raveled index -> z, y, x) takes an fp64-reciprocal path for volumes of more than 2^32 voxels;
tests/emul/coords_check.cpp compares it with integer division on random indices and around every plane boundary
of awkward shapes (host build of the same header the kernels use)."""
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)


def test_coords_fast_path_matches_integer_division(tmp_path):
    exe = str(tmp_path / 'coords_check')
    subprocess.check_call(['g++', '-O2', '-I', os.path.join(ROOT, 'skan_b200', 'csrc'), '-o', exe,
                           os.path.join(HERE, 'emul', 'coords_check.cpp')])
    out = subprocess.run([exe], capture_output=True, text=True)
    assert out.returncode == 0, out.stdout
    assert ' 0 bad' in out.stdout
