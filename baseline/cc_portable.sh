#!/bin/bash
# Compiler wrapper used only while pip-installing the unmodified reference into baseline/_ref: the reference's setup.py
# hard-codes -march=native (setup.py:176), but the binaries are built in the CPU-only container and run on the GPU
# box's host CPU, so the ISA level is pinned to x86-64-v3 instead (same choice as oracle/build_ref.py).
args=()
for a in "$@"; do
  if [ "$a" = "-march=native" ]; then args+=("-march=x86-64-v3"); else args+=("$a"); fi
done
exec /usr/bin/gcc "${args[@]}"
