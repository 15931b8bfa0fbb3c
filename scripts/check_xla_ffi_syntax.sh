#!/usr/bin/env bash
# SYNTAX / SIGNATURE check of the XLA-FFI shim (fortuna_b200/xla_ffi/fb200_xla_ffi.cc) where jaxlib is not installable:
# g++ -fsyntax-only against the MOCK header fortuna_b200/xla_ffi/mock/xla/ffi/api/ffi.h, whose Binding::To()
# static_asserts that every handler is callable with the argument list of its binding, and against the real
# include/fortuna_b200.h (so a drifted C-ABI signature fails too).  This is NOT an integration test: nothing is linked,
# registered with XLA or run.  With jaxlib installed, pass its include directory as $1 to check against the real header.
set -euo pipefail
here="$(cd "$(dirname "${BASH_SOURCE[0]}")/.." && pwd)"
inc="${1:-$here/fortuna_b200/xla_ffi/mock}"
cuda_inc="${CUDA_HOME:-/usr/local/cuda}/include"
g++ -std=c++17 -fsyntax-only -Wall -Wextra -Wno-unused-parameter \
    -I"$inc" -I"$here/include" -I"$cuda_inc" "$here/fortuna_b200/xla_ffi/fb200_xla_ffi.cc"
n=$(grep -c '^XLA_FFI_DEFINE_HANDLER_SYMBOL' "$here/fortuna_b200/xla_ffi/fb200_xla_ffi.cc")
echo "fb200_xla_ffi.cc: syntax ok against $inc ($n handlers)"
