"""python -m balm_b200.benchmark_virtual [--winSize 20 --sufSize 150 --ptsSize 40 --point_noise 0.05 --surf_range 2]

Headless run of the reference's synthetic benchmark (src/benchmark/benchmark_virtual.cpp:486-640; parameter names
and code defaults of :535-542, launch/benchmark_virtual.launch overrides winSize/sufSize to 20/20)."""
import argparse
import sys

from . import _lib as L
from . import drivers


def main(argv=None):
    ap = argparse.ArgumentParser(description=__doc__.splitlines()[0])
    ap.add_argument("--winSize", type=int, default=20)
    ap.add_argument("--sufSize", type=int, default=150)
    ap.add_argument("--ptsSize", type=int, default=40)
    ap.add_argument("--point_noise", type=float, default=0.05)
    ap.add_argument("--surf_range", type=float, default=2.0)
    ap.add_argument("--seed", type=int, default=10, help="replaces the reference's time(0) seed")
    ap.add_argument("--device", type=int, default=0)
    ap.add_argument("--precision", choices=["tensor", "fp64"], default="tensor")
    a = ap.parse_args(argv)
    drivers.benchmark_virtual(a.winSize, a.sufSize, a.ptsSize, a.point_noise, a.surf_range, a.seed, a.device,
                              L.PREC_TENSOR if a.precision == "tensor" else L.PREC_FP64)
    return 0


if __name__ == "__main__":
    sys.exit(main())
