for lib in "$@"; do D4B200_LIBRARY=$PWD/$lib python -m pytest tests/test_gpu_large.py -m gpu -q -x 2>&1 | tail -1; done
C4_AB_OUT=r02_ab_c4_dual.txt bash tools/c4_ab.sh "$@"
