set -x
tools/gpu_ab.sh r2i default b3:check w7:check default b3 w7
