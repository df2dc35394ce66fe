#!/bin/bash
# tests/hostemu/build.sh -- TEST TOOLING ONLY: host build of the encoder stage functions.
set -e
HERE="$(cd "$(dirname "$0")" && pwd)"
g++ -O2 -g -std=c++17 -fPIC -shared -Wall -Wno-unused-function -Wno-sign-compare -ffp-contract=off \
	"$HERE/hostemu.cpp" -o "$HERE/libhostemu.so"
