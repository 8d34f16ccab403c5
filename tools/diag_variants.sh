#!/bin/bash
# development: ratio (and speed) of the zstd compressor for prebuilt library variants (tools/bin/variants/lib_*.so)
cp znippy_b200/libznippy_cuda.so /tmp/lib_orig.so
for f in tools/bin/variants/lib_*.so; do
  cp $f znippy_b200/libznippy_cuda.so
  echo "== $f"; python tools/diag_compress.py 2>&1 | head -3 | cut -c1-60
done
cp /tmp/lib_orig.so znippy_b200/libznippy_cuda.so
