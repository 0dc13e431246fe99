"""Builds egonerf_b200/libegn_b200.so in-tree with nvcc for sm_100a (cross-compiles without a GPU)."""
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))


def build(verbose=False, force=False):
    csrc = os.path.join(HERE, "csrc")
    if force:
        subprocess.run(["make", "-C", csrc, "clean"], check=True, capture_output=not verbose)
    res = subprocess.run(["make", "-C", csrc, "-j8"], capture_output=True, text=True)
    if verbose or res.returncode != 0:
        print(res.stdout[-4000:])
        print(res.stderr[-4000:])
    if res.returncode != 0:
        raise RuntimeError("nvcc build of libegn_b200.so failed")
    return os.path.join(HERE, "libegn_b200.so")


if __name__ == "__main__":
    print(build(verbose=True))
