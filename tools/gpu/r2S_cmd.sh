mkdir -p gpurun_out
run() { echo "== $1 chains=$2"; JELLYFYSH_B200_LIBRARY=$PWD/build_variants/$1.so timeout 200 python tools/probe_water.py 32 $2 2000 2>&1 | grep -E "step [2]|rror" | cut -c1-110; }
run w8r120 2368; run w8r120 4096
run w8r104 2368
run w6r112 2664; run w6r112 4096
run w10r96 2960
run w4r120 2368; run w4r120 4096
