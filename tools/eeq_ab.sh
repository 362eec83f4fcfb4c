for lib in "$@"; do echo "== $lib"; D4B200_LIBRARY=$PWD/$lib python tools/eeq_time.py 2>&1 | tail -3; done | tee gpurun_out/r02_eeq_ab.txt
