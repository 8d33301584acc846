#!/bin/bash
# Sanitizer passes over the host-side replacements (CPU only; the harness pass needs oracle/_ref/obj_v3 = /root/reference built).
set -e
T=${TMPDIR:-/tmp}
g++ -std=c++17 -O1 -g -fsanitize=address,undefined -fno-sanitize-recover=undefined tests/pass_alive_replay.cc -o $T/replay_asan
python -c "import gzip; open('$T/pa_cases.bin','wb').write(gzip.open('tests/golden/pass_alive_cases.bin.gz','rb').read())"
$T/replay_asan $T/pa_cases.bin
if [ -d oracle/_ref/obj_v3 ]; then
  OBJS=$(find oracle/_ref/obj_v3 -name '*.o' | grep -v "/main.o" | grep -v "_weak.o")
  /usr/bin/g++ -std=c++17 -O1 -g -fsanitize=address,undefined -fno-sanitize-recover=undefined -DNDEBUG -DUSE_BLAS -DUSE_EIGEN -pthread -w \
      -march=x86-64-v3 -I/root/reference/src -I/root/reference/third_party/Eigen oracle/pass_alive_harness.cc $OBJS -o $T/harness_asan
  $T/harness_asan check 63 99
fi
g++ -std=c++17 -O1 -g -fsanitize=thread -pthread -Isayuri_b200/csrc/shim oracle/cache_harness.cc -o $T/cache_tsan
$T/cache_tsan bench 16 2 4096 20000 8
$T/cache_tsan bench 8 1 64 300 2
