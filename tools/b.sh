#!/bin/bash
# build the product library; exit non-zero (and show errors) on failure
cd "$(dirname "$0")/../arendur_b200/csrc" && make -s "$@" > /tmp/arn_make.log 2>&1 || { grep -E "error" build.log | head -20; echo BUILD FAILED; exit 1; }
grep -E "Compiling entry|Used [0-9]+ registers" build.log | sed 's/ptxas info    : //' | paste - - | sed -E "s/Compiling entry function '_ZN3arn[0-9]+([a-z_]+).*' for 'sm_100a'/\1/" | cut -c1-110 | grep -E "shade|connect|extend|trace|accum" 
