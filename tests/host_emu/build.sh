#!/bin/sh
# Build the host emulation of the tile functions (test infrastructure, see emu.cpp).
set -e
cd "$(dirname "$0")"
/usr/bin/g++ -std=c++17 -O2 -ffp-contract=off -fno-fast-math -fPIC -shared -o libpysp_emu.so emu.cpp
