#!/bin/bash
# AddressSanitizer and ThreadSanitizer runs of the host inflate (csrc/h_inflate.cu, several decoders on one stream).
# No GPU needed.  Usage: tools/sanitize_inflate/run.sh   (from the repo root; builds into /tmp/llfe_sanitize)
set -e
ROOT=$(cd "$(dirname "$0")/../.." && pwd)
W=/tmp/llfe_sanitize
mkdir -p $W
INC="-I $ROOT/include -I $ROOT/low_level_feature_extraction_b200/csrc"
SRC="$ROOT/low_level_feature_extraction_b200/csrc/h_inflate.cu $ROOT/tools/sanitize_inflate/stub.cu"
nvcc -O1 -g -std=c++17 -Wno-deprecated-gpu-targets -Xcompiler -fPIC,-fsanitize=address,-fno-omit-frame-pointer $INC -shared -o $W/libinf_asan.so $SRC -cudart static -Xlinker -lasan
sed "s#/tmp/asan/libinf_asan.so#$W/libinf_asan.so#" $ROOT/tools/sanitize_inflate/run_asan.py > $W/run_asan.py
LD_PRELOAD=$(gcc -print-file-name=libasan.so) ASAN_OPTIONS=detect_leaks=0 python $W/run_asan.py
python - <<PY
import zlib, numpy as np
rng = np.random.default_rng(3)
walk = (np.cumsum(rng.integers(-2, 3, 2_500_000)) & 255).astype(np.uint8).tobytes()
z = zlib.compress(walk, 6)
open("$W/z.bin", "wb").write(z); open("$W/ref.bin", "wb").write(walk)
b = bytearray(z); b[len(b) // 2] ^= 4
open("$W/zbad.bin", "wb").write(bytes(b))
PY
sed "s#/tmp/asan/#$W/#g" $ROOT/tools/sanitize_inflate/tsan_driver.cu > $W/tsan_driver.cu
nvcc -O1 -g -std=c++17 -Wno-deprecated-gpu-targets -Xcompiler -fsanitize=thread,-fno-omit-frame-pointer $INC -o $W/drv_tsan $W/tsan_driver.cu $SRC -cudart static -Xlinker -ltsan
$W/drv_tsan
