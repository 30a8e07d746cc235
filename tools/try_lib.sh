#!/bin/bash
# usage: tools/try_lib.sh <variant.so> <cmd...>  — runs cmd with the variant library in place of libarn_b200.so
set -e
cp arendur_b200/libarn_b200.so /tmp/libarn_saved.so
cp "$1" arendur_b200/libarn_b200.so; shift
"$@" || true
cp /tmp/libarn_saved.so arendur_b200/libarn_b200.so
