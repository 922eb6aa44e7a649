#!/bin/bash
# build the library; print BUILD OK / BUILD FAILED (+ the first errors) and return the status
cd /root/repo || exit 1
out=$(python nmfk.jl_b200/build.py 2>&1); rc=$?
if [ $rc -ne 0 ]; then echo "$out" | grep -E "error" | head -8; echo "BUILD FAILED"; exit 1; fi
echo "$out" | tail -1; echo "BUILD OK"
