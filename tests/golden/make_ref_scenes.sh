#!/bin/bash
# Regenerates tests/golden/ref_scenes/*.bin: the geometry of the reference's own tests/test00, test01, test03 clients exactly
# as AcceleratorB200 extracts and uploads it (B200_DUMP_SCENE, integration/src/accelerator/accelerator_b200.cc).  Build
# container only (needs /root/reference and `make -C integration`); no GPU needed -- the dump is written before libb200rt is
# asked for a device.  tests/test02 (396 830 primitives, 21 MB) is dumped the same way for tests/tools/dump_compare.py but not
# committed; scenes.cube_grid reproduces what made it special.
set -e
root=$(cd "$(dirname "$0")/../.." && pwd)
out=$root/tests/golden/ref_scenes
mkdir -p "$out"
tmp=$(mktemp -d)
for t in test00 test01 test03; do
  bin=$root/integration/_build/yafaray_$t
  if [ ! -x "$bin" ]; then
    gcc -O2 -w -I$root/integration/ref_shims -I/root/reference/include/public_api -include $root/integration/test_hook.h \
        /root/reference/tests/$t/$t.c -o $tmp/yafaray_$t -L$root/integration/_build -lyafaray4_b200 \
        -Wl,-rpath,$root/integration/_build -Wl,-rpath,$root/libyafaray_b200 -lm
    bin=$tmp/yafaray_$t
  fi
  (cd $tmp && B200_AA_PASSES=1 B200_DETERMINISTIC=1 B200_ACCEL_TYPE=b200-kdtree B200_VERIFY_EXTRACTION=1 B200_DUMP_SCENE=$out/$t.bin $bin 2>&1 | grep -a "extraction check" | tail -1)
done
ls -la "$out"
